// Kernels of the shooting loop (sm_100a).  One shot of the reference's parallelRadiosity()
// (src/rad.cpp:406-451) is:
//
//   K1+K2  select_kernel / the fused tail of update_select_kernel: argmax of unshot energy over the
//          receivers (warp shuffle -> block -> last-block-done), then one thread derives the
//          shooter's two half samples (listCanditate rad.cpp:387-404, AdaptCalFF/PartitionSource
//          shade.cpp:146-161, 76-99).
//   K3a    vertex_ff_kernel: one thread per (unique receiver vertex, shooter half): facing culls,
//          analytic form factor, visibility walk through the flattened octree (CalFF shade.cpp:107-127,
//          RayHitted raycast.cpp:335-360).
//   K3b    update_select_kernel: one thread per receiver: gathers its three vertex form factors,
//          applies CalRadiosity (shade.cpp:244-259) to accB/deltaB, and feeds the next shot's argmax.
//   (generic) shade_receivers_kernel: per-receiver evaluation without vertex sharing, used for
//          -parallel_src > 1 and when vertex de-duplication is switched off.
//
// HBM layout: structure of float4 planes (one plane per triangle vertex), receivers sorted by
// (original parent, Morton code) so a warp's 32 receivers are neighbours in space:
//   geo[v][r] = p[v].xyz | v=0: parent, v=1: slot, v=2: area
//   st[v][r]  = accB[v].rgb | deltaB[v]            <- the only read-modify-write stream
//   vid[r]    = unique-vertex ids of the three corners | parent;  slot[r] = TriStore slot
//   uv[u]     = unique vertex position | parent
// Per-original tables (n, |n|, Frgb*k) and the octree stay L2/L1 resident.
#pragma once
#include "rt_core.cuh"
#include "shot_core.cuh"

namespace radk {

constexpr int kMaxSources = 64;		// upper bound on -parallel_src handled on the device

struct DevStats {	// device-side counters, unsigned long long each (mirrors radk_stats)
	unsigned long long shots, rays, visible, calff, traced, leaf_visits, descent_steps, tri_tests, tri_accepts,
		precision_exceeded, receivers, splits, guard_trips;
};

struct SelectState {
	unsigned long long *partial_key;	// [max blocks]
	unsigned *partial_cnt;			// [max blocks] candidates seen by the block
	unsigned *ticket;			// last-block-done counter
	unsigned long long *bound;		// previous winner's key for multi-source selection (exclusive)
	int *ncand;				// |C| of the current selection (device)
	int *nsrc;				// number of shooters chosen for the current iteration
	int *converged;				// set when a selection finds no candidate
	long long *shots_done;			// shots applied so far
	int *shot_log;				// slot of the first shooter of each shot (ring of shot_log_cap)
	int shot_log_cap;
};

struct DevScene {
	DevTree tree;
	int R;				// receivers (leaf slots) on this shard
	int U;				// unique (vertex, parent) pairs on this shard
	int n0;
	const float4 *geo[3];
	float4 *st[3];
	const int4 *vid;
	const int *slot;		// [R] TriStore slot of receiver r (tie-break key, output order)
	const float4 *uv;
	const int *uv_leaf;		// [U] leaf containing the vertex (TreeNodeNum of the ray's first probe point, static)
	float2 *ffh;			// [U] form factor toward half 0 / half 1 of the current shooter
	unsigned *vstat;		// [U] statistics mode: packed sample counts
	const float4 *par_n;		// [n0] n.xyz | norm(n)
	const float4 *par_fk;		// [n0] (float)Frgb[c] * (float)(0.5/255.0)
	const int *slot_to_recv;	// [trinum] receiver index of a slot on this shard, -1 otherwise
	float converge;
	int parallel_src;
	DevStats *stats;
	SelectState sel;
};

// ---- selection -------------------------------------------------------------------------------------
// keys: shot_core.cuh (make_key / energy_key)
__device__ __forceinline__ float unordered_bits(unsigned o) {
	return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// warp-wide maximum of 64-bit keys with two redux.sync instructions: maximum of the high words, then
// maximum of the low words among the lanes that hold that high word
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k) {
	const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
	const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
	const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
	return ((unsigned long long)mh << 32) | ml;
}

// K2: derive the shooter record of receiver r (one thread).
__device__ void build_shooter(const DevScene &S, int r, ShooterRec *out) {
	const float4 g0 = S.geo[0][r], g1 = S.geo[1][r], g2 = S.geo[2][r];
	const float p[3][3] = {{g0.x, g0.y, g0.z}, {g1.x, g1.y, g1.z}, {g2.x, g2.y, g2.z}};
	const int parent = __float_as_int(g0.w);
	const float4 pn = S.par_n[parent];
	make_shooter(p, pn.x, pn.y, pn.z, pn.w, g2.w, S.st[0][r].w, S.st[1][r].w, S.st[2][r].w, parent,
		     __float_as_int(g1.w), r, out);
}

// Block-level part of the argmax; returns true in thread 0 of the LAST block to finish, with the
// grid-wide maximum in *key_out and the candidate count in *cnt_out.
__device__ bool grid_argmax(const SelectState &sel, unsigned long long key, unsigned is_cand,
			    unsigned long long *key_out, unsigned *cnt_out) {
	__shared__ unsigned long long s_key[32];
	__shared__ unsigned s_cnt[32];
	__shared__ bool s_last;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
	key = warp_max_u64(key);
	unsigned cnt = __reduce_add_sync(0xffffffffu, is_cand);
	if (lane == 0) s_key[warp] = key, s_cnt[warp] = cnt;
	__syncthreads();
	if (warp == 0) {
		key = lane < nwarps ? s_key[lane] : 0ull;
		cnt = lane < nwarps ? s_cnt[lane] : 0u;
		key = warp_max_u64(key);
		cnt = __reduce_add_sync(0xffffffffu, cnt);
		if (lane == 0) {
			sel.partial_key[blockIdx.x] = key;
			sel.partial_cnt[blockIdx.x] = cnt;
			__threadfence();
			const unsigned done = atomicAdd(sel.ticket, 1u);
			s_last = (done == gridDim.x - 1);
		}
	}
	__syncthreads();
	if (!s_last) return false;
	__threadfence();
	unsigned long long best = 0ull;
	unsigned total = 0u;
	for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
		const unsigned long long k = __ldcg(&sel.partial_key[b]);
		best = k > best ? k : best;
		total += __ldcg(&sel.partial_cnt[b]);
	}
	best = warp_max_u64(best);
	total = __reduce_add_sync(0xffffffffu, total);
	__syncthreads();
	if (lane == 0) s_key[warp] = best, s_cnt[warp] = total;
	__syncthreads();
	if (warp == 0) {
		best = lane < nwarps ? s_key[lane] : 0ull;
		total = lane < nwarps ? s_cnt[lane] : 0u;
		best = warp_max_u64(best);
		total = __reduce_add_sync(0xffffffffu, total);
		if (lane == 0) {
			*key_out = best, *cnt_out = total;
			*sel.ticket = 0u;
			return true;
		}
	}
	return false;
}

// Candidate key of receiver r given its (possibly just updated) deltaB.
__device__ __forceinline__ unsigned long long candidate_key(const DevScene &S, float d0, float d1, float d2, int slot,
							     unsigned long long bound, unsigned *is_cand) {
	const float delta = d0 + d1 + d2;
	*is_cand = 0u;
	if (!(delta > S.converge)) return 0ull;
	*is_cand = 1u;
	const unsigned long long k = make_key(delta, slot);
	return k < bound ? k : 0ull;
}

// Stand-alone K1 (first shot, multi-source passes, API select).  `which` = index of the source being
// chosen in this iteration; pass 0 also publishes |C| and the number of sources n = min((|C|+4)/5, P).
__global__ void __launch_bounds__(256) select_kernel(DevScene S, ShooterRec *out, int which) {
	const unsigned long long bound = which == 0 ? ~0ull : *S.sel.bound;
	if (which > 0 && which >= *S.sel.nsrc) {	// fewer sources than passes: nothing to do
		if (blockIdx.x == 0 && threadIdx.x == 0) out->valid = 0;
		return;
	}
	unsigned long long key = 0ull;
	unsigned cand = 0u;
	for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < S.R; r += gridDim.x * blockDim.x) {
		unsigned c;
		const unsigned long long k = candidate_key(S, S.st[0][r].w, S.st[1][r].w, S.st[2][r].w,
							   __ldg(&S.slot[r]), bound, &c);
		key = k > key ? k : key;
		cand += c;
	}
	unsigned long long best;
	unsigned total;
	if (!grid_argmax(S.sel, key, cand, &best, &total)) return;
	if (which == 0) {
		int n = ((int)total + 4) / 5;
		n = n < S.parallel_src ? n : S.parallel_src;
		*S.sel.ncand = (int)total;
		*S.sel.nsrc = n;
		*S.sel.converged = (total == 0u);
	}
	if (best == 0ull) { out->valid = 0; return; }
	*S.sel.bound = best;
	build_shooter(S, S.slot_to_recv[(int)(best & 0xffffffffu)], out);
}

// ---- statistics plumbing -----------------------------------------------------------------------
__device__ __forceinline__ void flush_counters(DevStats *st, const FFCounters &fc, const WalkCounters &wc, unsigned recv) {
	unsigned v[11] = {fc.rays, fc.visible, fc.calff, wc.traced, wc.leaf_visits, wc.descent_steps, wc.tri_tests,
			  wc.tri_accepts, wc.precision_exceeded, recv, wc.guard_trips};
	unsigned long long *dst[11] = {&st->rays, &st->visible, &st->calff, &st->traced, &st->leaf_visits, &st->descent_steps,
				       &st->tri_tests, &st->tri_accepts, &st->precision_exceeded, &st->receivers, &st->guard_trips};
#pragma unroll
	for (int i = 0; i < 11; i++) {
		const unsigned s = __reduce_add_sync(0xffffffffu, v[i]);
		if ((threadIdx.x & 31) == 0 && s) atomicAdd(dst[i], (unsigned long long)s);
	}
}

// ---- K3a: form factors of the unique receiver vertices -------------------------------------------
// thread t < U handles half 0, t >= U half 1: a warp's rays share one target point and start from
// neighbouring vertices, so they walk the same cells.
template <bool STATS>
__global__ void __launch_bounds__(128) vertex_ff_kernel(DevScene S, const ShooterRec *__restrict__ shp) {
	__shared__ ShooterRec sh;
	if (threadIdx.x < sizeof(ShooterRec) / 4) ((int *)&sh)[threadIdx.x] = ((const int *)shp)[threadIdx.x];
	__syncthreads();
	if (!sh.valid || *S.sel.converged) return;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int nthreads = sh.splittable ? 2 * S.U : S.U;
	FFCounters fc = {0, 0, 0};
	WalkCounters wc = {0, 0, 0, 0, 0, 0, 0};
	if (t < nthreads) {
		const int half = t >= S.U;
		const int u = half ? t - S.U : t;
		const float4 v = __ldg(&S.uv[u]);
		const int parent = __float_as_int(v.w);
		const float4 pn = __ldg(&S.par_n[parent]);
		const FFSample &smp = sh.splittable ? sh.half[half] : sh.centre;
		const float ff = cal_ff<STATS>(S.tree, smp, sh.logsrc, pn.x, pn.y, pn.z, pn.w, parent, v.x, v.y, v.z, &fc, &wc);
		if (sh.splittable) {
			if (half) S.ffh[u].y = ff; else S.ffh[u].x = ff;
		} else {
			S.ffh[u] = make_float2(ff, 0.0f);
		}
		if (STATS) {
			// reference-equivalent sample counts of this vertex: bits 0-1 rays, 2-3 visible (this half),
			// the discarded centre sample is accounted for by half 0's thread
			unsigned packed = fc.rays | (fc.visible << 2);
			if (sh.splittable && !half) {
				FFCounters cc = {0, 0, 0};
				const float dx = sh.centre.cx - v.x, dy = sh.centre.cy - v.y, dz = sh.centre.cz - v.z;
				const float dotSrc = dot3(dx, dy, dz, sh.centre.nx, sh.centre.ny, sh.centre.nz);
				const float dotDes = -dot3(dx, dy, dz, pn.x, pn.y, pn.z);
				if ((dotSrc < 0) == (dotDes < 0)) {
					const float n2 = dx * dx + dy * dy + dz * dz;
					const float dirLen = sqrtf(n2);
					float f = (dotSrc / dirLen / sh.centre.nlen) * (dotDes / dirLen / pn.w);
					if (!(f <= 0.0f)) {
						f = (float)((double)f * ((double)sh.centre.area / ((double)n2 * 3.1415926535897932)));
						if (!(f <= 0.0f)) {
							cc.rays = 1;
							cc.visible = ray_walk<false>(S.tree, v.x, v.y, v.z, dx, dy, dz, parent, nullptr) == sh.logsrc;
						}
					}
				}
				packed += cc.rays | (cc.visible << 2);
			}
			// two halves write disjoint bytes
			((unsigned char *)&S.vstat[u])[half] = (unsigned char)packed;
		}
	}
	if (STATS) {
		FFCounters none = {0, 0, 0};	// rays/visible/calff are counted per receiver in K3b
		flush_counters(S.stats, none, wc, 0);
	}
}

// ---- K3b: radiosity update of the receivers + next shot's argmax --------------------------------
template <bool STATS>
__global__ void __launch_bounds__(256) update_select_kernel(DevScene S, const ShooterRec *__restrict__ shp, ShooterRec *next) {
	__shared__ ShooterRec sh;
	if (threadIdx.x < sizeof(ShooterRec) / 4) ((int *)&sh)[threadIdx.x] = ((const int *)shp)[threadIdx.x];
	__syncthreads();
	if (!sh.valid || *S.sel.converged) return;
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long key = 0ull;
	unsigned cand = 0u, shaded = 0u;
	FFCounters fc = {0, 0, 0};
	if (r < S.R) {
		float4 a0 = S.st[0][r], a1 = S.st[1][r], a2 = S.st[2][r];
		const int4 id = __ldg(&S.vid[r]);
		const int slot = __ldg(&S.slot[r]);
		if (slot == sh.slot) {
			// the shooter does not receive from itself; its unshot energy is now spent (rad.cpp:446-449)
			a0.w = a1.w = a2.w = 0.0f;
			S.st[0][r].w = 0.0f, S.st[1][r].w = 0.0f, S.st[2][r].w = 0.0f;
		} else if (a0.x < 255.0f || a0.y < 255.0f || a0.z < 255.0f || a1.x < 255.0f || a1.y < 255.0f || a1.z < 255.0f ||
			   a2.x < 255.0f || a2.y < 255.0f || a2.z < 255.0f) {	// rad.cpp:424-428
			shaded = 1u;
			const float2 h0 = S.ffh[id.x], h1 = S.ffh[id.y], h2 = S.ffh[id.z];
			const float ff[3] = {h0.x + h0.y, h1.x + h1.y, h2.x + h2.y};
			const float4 fk = __ldg(&S.par_fk[id.w]);
			// CalRadiosity (shade.cpp:244-259): lo[c] = Frgb[c]*k*deltaB_src[c]/3
			const float lo[3] = {fk.x * sh.dB[0] / 3, fk.y * sh.dB[1] / 3, fk.z * sh.dB[2] / 3};
			float dB[3] = {a0.w, a1.w, a2.w};
			float4 *acc[3] = {&a0, &a1, &a2};
#pragma unroll
			for (int v = 0; v < 3; v++) {
				float d;
				d = lo[0] * ff[v]; dB[0] += d; acc[v]->x += d;
				d = lo[1] * ff[v]; dB[1] += d; acc[v]->y += d;
				d = lo[2] * ff[v]; dB[2] += d; acc[v]->z += d;
			}
			a0.w = dB[0], a1.w = dB[1], a2.w = dB[2];
			S.st[0][r] = a0, S.st[1][r] = a1, S.st[2][r] = a2;
			if (STATS) {
				const unsigned s0 = S.vstat[id.x], s1 = S.vstat[id.y], s2 = S.vstat[id.z];
				const unsigned b = (s0 & 0xff) + ((s0 >> 8) & 0xff), c = (s1 & 0xff) + ((s1 >> 8) & 0xff),
					       e = (s2 & 0xff) + ((s2 >> 8) & 0xff);
				fc.rays = (b & 3) + (c & 3) + (e & 3);
				fc.visible = ((b >> 2) & 3) + ((c >> 2) & 3) + ((e >> 2) & 3);
				fc.calff = sh.splittable ? 9 : 3;
			}
		}
		key = candidate_key(S, a0.w, a1.w, a2.w, slot, ~0ull, &cand);
	}
	if (STATS) {
		WalkCounters none = {0, 0, 0, 0, 0, 0, 0};
		flush_counters(S.stats, fc, none, shaded);
	}
	unsigned long long best;
	unsigned total;
	if (!grid_argmax(S.sel, key, cand, &best, &total)) return;
	// tail executed by one thread of the last block: bookkeeping of the shot just applied + K2 of the next
	const long long k = (*S.sel.shots_done)++;
	S.sel.shot_log[k % S.sel.shot_log_cap] = sh.slot;
	atomicAdd(&S.stats->shots, 1ull);
	*S.sel.ncand = (int)total;
	if (best == 0ull) {
		// converged: no record may stay valid, later launches of the same batch must be no-ops
		next->valid = 0;
		const_cast<ShooterRec *>(shp)->valid = 0;
		*S.sel.converged = 1;
		return;
	}
	build_shooter(S, S.slot_to_recv[(int)(best & 0xffffffffu)], next);
}

// ---- generic per-receiver shading (multi-source / no vertex sharing) ----------------------------
// Restates the body of the receiver loop (rad.cpp:417-440): the saturation test is evaluated once,
// then the sources are applied in selection order.
template <bool STATS>
__global__ void __launch_bounds__(128) shade_receivers_kernel(DevScene S, const ShooterRec *__restrict__ list, int list_cap) {
	__shared__ ShooterRec sh[8];
	const int nsrc_all = min(*S.sel.nsrc, list_cap);
	if (nsrc_all <= 0) return;
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	FFCounters fc = {0, 0, 0};
	WalkCounters wc = {0, 0, 0, 0, 0, 0, 0};
	unsigned shaded = 0u;
	float4 a0, a1, a2, g0, g1, g2, pn, fk;
	bool active = false, is_src = false;
	int parent = 0, slot = -1;
	if (r < S.R) {
		a0 = S.st[0][r], a1 = S.st[1][r], a2 = S.st[2][r];
		g0 = __ldg(&S.geo[0][r]), g1 = __ldg(&S.geo[1][r]), g2 = __ldg(&S.geo[2][r]);
		parent = __float_as_int(g0.w), slot = __float_as_int(g1.w);
		pn = __ldg(&S.par_n[parent]), fk = __ldg(&S.par_fk[parent]);
		active = (a0.x < 255.0f || a0.y < 255.0f || a0.z < 255.0f || a1.x < 255.0f || a1.y < 255.0f || a1.z < 255.0f ||
			  a2.x < 255.0f || a2.y < 255.0f || a2.z < 255.0f);
	}
	for (int base = 0; base < nsrc_all; base += 8) {	// sources staged through shared memory, 8 at a time
		const int nb = min(8, nsrc_all - base);
		__syncthreads();
		for (int i = threadIdx.x; i < nb * (int)(sizeof(ShooterRec) / 4); i += blockDim.x)
			((int *)sh)[i] = ((const int *)(list + base))[i];
		__syncthreads();
		for (int j = 0; j < nb; j++) is_src |= (sh[j].valid && sh[j].slot == slot);
	}
	if (r < S.R && active && !is_src) shaded = 1u;
	for (int base = 0; base < nsrc_all; base += 8) {
		const int nb = min(8, nsrc_all - base);
		__syncthreads();
		for (int i = threadIdx.x; i < nb * (int)(sizeof(ShooterRec) / 4); i += blockDim.x)
			((int *)sh)[i] = ((const int *)(list + base))[i];
		__syncthreads();
		if (!shaded) continue;
		for (int j = 0; j < nb; j++) {
			if (!sh[j].valid) continue;
			const float4 *g[3] = {&g0, &g1, &g2};
			float ff[3];
			for (int v = 0; v < 3; v++)
				ff[v] = adapt_ff<STATS>(S.tree, sh[j], pn.x, pn.y, pn.z, pn.w, parent, g[v]->x, g[v]->y, g[v]->z, &fc, &wc);
			const float lo[3] = {fk.x * sh[j].dB[0] / 3, fk.y * sh[j].dB[1] / 3, fk.z * sh[j].dB[2] / 3};
			float4 *acc[3] = {&a0, &a1, &a2};
			float dB[3] = {a0.w, a1.w, a2.w};
#pragma unroll
			for (int v = 0; v < 3; v++) {
				float d;
				d = lo[0] * ff[v]; dB[0] += d; acc[v]->x += d;
				d = lo[1] * ff[v]; dB[1] += d; acc[v]->y += d;
				d = lo[2] * ff[v]; dB[2] += d; acc[v]->z += d;
			}
			a0.w = dB[0], a1.w = dB[1], a2.w = dB[2];
		}
	}
	if (r < S.R) {
		if (is_src) a0.w = a1.w = a2.w = 0.0f;	// rad.cpp:446-449
		if (shaded || is_src) S.st[0][r] = a0, S.st[1][r] = a1, S.st[2][r] = a2;
	}
	if (STATS) flush_counters(S.stats, fc, wc, shaded);
}

// bookkeeping after a generic pass (one thread)
__global__ void after_generic_pass_kernel(DevScene S, const ShooterRec *list) {
	if (*S.sel.nsrc <= 0) return;
	const long long k = (*S.sel.shots_done)++;
	S.sel.shot_log[k % S.sel.shot_log_cap] = list[0].slot;
	atomicAdd(&S.stats->shots, 1ull);
}

// caller-chosen shooters (radk_ctx_shoot / radk_vertex_ff): K2 for each listed receiver
__global__ void make_sources_kernel(DevScene S, const int *recv, int n, ShooterRec *list, int set_nsrc) {
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < n) {
		if (recv[j] >= 0) build_shooter(S, recv[j], &list[j]);
		else list[j].valid = 0;
	}
	if (j == 0 && set_nsrc) *S.sel.nsrc = n;
}

// ---- test / utility kernels ----------------------------------------------------------------------
__global__ void ray_batch_kernel(DevTree T, const float *p, const float *v, const int *from, int n, int *hit) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	hit[i] = ray_walk<false>(T, p[3 * i], p[3 * i + 1], p[3 * i + 2], v[3 * i], v[3 * i + 1], v[3 * i + 2], from[i], nullptr);
}

// radk_ray_sees: leaf mask of one original (bit n <=> leaf n lists it), then the any-hit walk of the batched path
__global__ void leaf_mask_kernel(DevTree T, int logsrc, unsigned *mask) {
	for (int k = T.tri_leaf_begin[logsrc] + threadIdx.x; k < T.tri_leaf_begin[logsrc + 1]; k += blockDim.x) {
		const int leaf = T.tri_leaves[k];
		atomicOr(&mask[leaf >> 5], 1u << (leaf & 31));
	}
}
__global__ void ray_sees_batch_kernel(DevTree T, const unsigned *mask, const float *p, const float *v, const int *from, int logsrc, int n, int *out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	out[i] = ray_sees(T, mask, p[3 * i], p[3 * i + 1], p[3 * i + 2], v[3 * i], v[3 * i + 1], v[3 * i + 2], from[i], logsrc) ? 1 : 0;
}

__global__ void vertex_ff_batch_kernel(DevScene S, const ShooterRec *shp, const int *recv, const int *vertex, int n, float *out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int r = recv[i];
	if (r < 0 || r >= S.R) { out[i] = 0.0f; return; }
	const float4 g = S.geo[vertex[i]][r];
	const int parent = __float_as_int(S.geo[0][r].w);
	const float4 pn = S.par_n[parent];
	FFCounters fc;
	WalkCounters wc;
	out[i] = adapt_ff<false>(S.tree, *shp, pn.x, pn.y, pn.z, pn.w, parent, g.x, g.y, g.z, &fc, &wc);
}

// K5: snapshot colour per receiver (PrintOut rad.cpp:168-178)
__global__ void snapshot_kernel(DevScene S, int *rgb_by_recv) {
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= S.R) return;
	const float4 a0 = S.st[0][r], a1 = S.st[1][r], a2 = S.st[2][r];
	const float s[3] = {a0.x + a1.x + a2.x, a0.y + a1.y + a2.y, a0.z + a1.z + a2.z};
#pragma unroll
	for (int j = 0; j < 3; j++) {
		float f = s[j] / 3;
		const float lo = (255.0f < f) ? 255.0f : f;
		f = (lo < 0.0f) ? 0.0f : lo;
		if (f != f) f = 0;
		rgb_by_recv[3 * r + j] = (int)f;
	}
}

}  // namespace radk
