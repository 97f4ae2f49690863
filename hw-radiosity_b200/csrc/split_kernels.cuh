// K4 - adaptive splitting inside the shooting loop (north_star subsystem 4; the reference's dormant
// path: Shade src/shade.cpp:463-525 with PartitionTriangleWithCenterAndStore 261-306 and the 3-way
// PartitionDestination 213-241, driven by the do/while of parallelRadiosity src/rad.cpp:414-444).
//
// The reference can only run this path serially (it races on trinum and rand()).
// Here one round is three data-parallel steps over the slots [lo, hi) of the pool:
//   split_shade_kernel   per slot: vertex form factors, the form-factor-gradient test (centre sample,
//                        then area/SampleArea random interior samples), decision; receivers that are
//                        not split get CalRadiosity at once, the others record (#new slots, ff, ffs);
//   split_scan_*         exclusive prefix sum of the #new slots (block sums -> one-CTA scan);
//   split_append_kernel  stream-compacted append: child records are written at pool[trinum + prefix],
//                        in slot order, exactly where the serial reference would have put them
//                        (3 new slots when an original becomes logical, 2 when a child is replaced in place).
// The appended range is the next round's [lo, hi); rounds repeat until the pool stops growing.
//
// Deviation from the reference (stated in DESIGN.md): the interior sample points come from a
// counter-based hash of (shot, slot, sample) instead of glibc rand(), whose sequence depends on the
// serial visiting order.  Split decisions can therefore differ near the DeltaFFLimit threshold;
// parity for this path is statistical (leaf count, energy), as SURVEY.md section 8c prescribes.
#pragma once
#include "shoot_kernels.cuh"

namespace radk {

struct SplitScene {
	DevTree tree;
	int capacity;			// pool slots allocated
	int limit;			// min(MaxTri, TriangleLimit) (shade.cpp:26-30)
	int *trinum;			// live slots (device)
	float4 *p[3];			// p[v].xyz | v=0: parent (int), v=1: area, v=2: -
	float4 *n;			// n.xyz | norm(n)
	float4 *st[3];			// accB[v].rgb | deltaB[v]
	float4 *fk;			// (float)Frgb*k rgb | -
	radk_triangle *rec;		// full records (static fields: Frgb, Brgb, se, neighbor, deltaaccB)
	int *count;			// [capacity] new slots requested by the slot this round (0, 2 or 3)
	float4 *ffs;			// [capacity] ff[0..2] | form factor at the centre
	int *block_sum, *block_off;	// scan scratch
	float converge, area_limit, sample_area, delta_ff;
	DevStats *stats;
	SelectState sel;
	unsigned long long shot_no;	// for the sample hash
};

__device__ __forceinline__ unsigned hash3(unsigned a, unsigned b, unsigned c) {
	unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
	h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
	return h;
}

__device__ void split_build_shooter(const SplitScene &S, int slot, ShooterRec *out) {
	const float4 g0 = S.p[0][slot], g1 = S.p[1][slot], g2 = S.p[2][slot];
	const float p[3][3] = {{g0.x, g0.y, g0.z}, {g1.x, g1.y, g1.z}, {g2.x, g2.y, g2.z}};
	const float4 pn = S.n[slot];
	make_shooter(p, pn.x, pn.y, pn.z, pn.w, g1.w, S.st[0][slot].w, S.st[1][slot].w, S.st[2][slot].w, __float_as_int(g0.w), slot,
		     slot, out);
}

// K1 over the pool (slots with parent >= 0)
__global__ void __launch_bounds__(256) split_select_kernel(SplitScene S, ShooterRec *out) {
	const int n = *S.trinum;
	unsigned long long key = 0ull;
	unsigned cand = 0u;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		if (__float_as_int(S.p[0][i].w) < 0) continue;
		const float delta = S.st[0][i].w + S.st[1][i].w + S.st[2][i].w;
		if (!(delta > S.converge)) continue;
		cand++;
		const unsigned long long k = make_key(delta, i);
		key = k > key ? k : key;
	}
	unsigned long long best;
	unsigned total;
	if (!grid_argmax(S.sel, key, cand, &best, &total)) return;
	*S.sel.ncand = (int)total;
	*S.sel.converged = (total == 0u);
	if (best == 0ull) { out->valid = 0; return; }
	split_build_shooter(S, (int)(best & 0xffffffffu), out);
}

__device__ __forceinline__ void split_cal_radiosity(const ShooterRec &sh, const float4 &fk, const float ff[3], float4 &a0, float4 &a1, float4 &a2) {
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
}

// Shade (shade.cpp:435-525) for the slots [lo, hi) of this round.  kSplitGroup lanes work on one slot and every
// form factor is evaluated by three of them (adapt_ff_sub: centre sample, half t1, half t2 - three independent octree
// walks): lanes 0-11 take the three vertices and the centre sample, then the random interior samples go five at a time.
// The serial loop stops at the first sample whose form factor leaves the plane by more than DeltaFFLimit; the group
// finds that sample with a ballot and only counts the rays the serial loop would have cast, so decisions, radiosity
// and statistics equal a one-thread-per-slot evaluation (one slot used to take up to (4 + area/SampleArea) x 3
// walks in sequence: 2.2 ms for the first round of a shot on blocks.tri).
constexpr int kSplitGroup = 16;
constexpr int kSplitBlock = 128;
constexpr int kSplitBatch = kSplitGroup / 3;	// interior samples per pass

__device__ __forceinline__ void add_counters(FFCounters &fc, WalkCounters &wc, const FFCounters &f, const WalkCounters &w) {
	fc.calff += f.calff, fc.rays += f.rays, fc.visible += f.visible;
	wc.traced += w.traced, wc.leaf_visits += w.leaf_visits, wc.descent_steps += w.descent_steps, wc.tri_tests += w.tri_tests;
	wc.tri_accepts += w.tri_accepts, wc.precision_exceeded += w.precision_exceeded, wc.guard_trips += w.guard_trips;
}

// value of point q (lanes 3q .. 3q+2 of the group hold its three parts)
__device__ __forceinline__ float group_ff(unsigned gmask, float part, int q, bool splittable) {
	const float v0 = __shfl_sync(gmask, part, 3 * q, kSplitGroup), v1 = __shfl_sync(gmask, part, 3 * q + 1, kSplitGroup),
		    v2 = __shfl_sync(gmask, part, 3 * q + 2, kSplitGroup);
	return splittable ? v1 + v2 : v0;
}

__global__ void __launch_bounds__(kSplitBlock) split_shade_kernel(SplitScene S, const ShooterRec *__restrict__ shp, int lo, int hi) {
	__shared__ ShooterRec sh;
	if (threadIdx.x < sizeof(ShooterRec) / 4) ((int *)&sh)[threadIdx.x] = ((const int *)shp)[threadIdx.x];
	__syncthreads();
	const int gl = threadIdx.x % kSplitGroup;	// lane within the slot's group
	const int i = lo + blockIdx.x * (kSplitBlock / kSplitGroup) + threadIdx.x / kSplitGroup;
	const int gshift = (threadIdx.x & 31) / kSplitGroup * kSplitGroup;
	const unsigned gmask = ((1u << kSplitGroup) - 1u) << gshift;
	const int q = gl / 3, sub = gl - 3 * q;		// point and part this lane evaluates (q < 5)
	FFCounters fc = {0, 0, 0};
	WalkCounters wc = {0, 0, 0, 0, 0, 0, 0};
	unsigned shaded = 0u;
	if (gl == 0 && i < hi) S.count[i] = 0;
	if (sh.valid && i < hi) {
		const float4 g0 = S.p[0][i], g1 = S.p[1][i], g2 = S.p[2][i];
		const int parent = __float_as_int(g0.w);
		float4 a0 = S.st[0][i], a1 = S.st[1][i], a2 = S.st[2][i];
		const bool unsat = a0.x < 255.0f || a0.y < 255.0f || a0.z < 255.0f || a1.x < 255.0f || a1.y < 255.0f || a1.z < 255.0f ||
				   a2.x < 255.0f || a2.y < 255.0f || a2.z < 255.0f;
		if (parent >= 0 && i != sh.slot && unsat) {	// rad.cpp:420-428
			shaded = gl == 0 ? 1u : 0u;
			const float4 pn = S.n[i];
			const float area = g1.w;
			const bool splittable = sh.splittable != 0;
			// shade.cpp:455-466: too small to split, or no room in the pool (re-checked exactly at append time)
			const bool try_split = !(area < S.area_limit) && !((double)area < 1e-8) && !(*S.trinum + 4 >= S.limit);
			// points 0-2: the vertices, point 3: the centre (only when a split is possible)
			const float cx = (float)((double)(g0.x + g1.x + g2.x) / 3.0), cy = (float)((double)(g0.y + g1.y + g2.y) / 3.0),
				    cz = (float)((double)(g0.z + g1.z + g2.z) / 3.0);
			float part = 0.0f;
			if (q < 3 || (q == 3 && try_split)) {
				const float qx = q == 0 ? g0.x : q == 1 ? g1.x : q == 2 ? g2.x : cx;
				const float qy = q == 0 ? g0.y : q == 1 ? g1.y : q == 2 ? g2.y : cy;
				const float qz = q == 0 ? g0.z : q == 1 ? g1.z : q == 2 ? g2.z : cz;
				part = adapt_ff_sub<true>(S.tree, sh, sub, pn.x, pn.y, pn.z, pn.w, parent, qx, qy, qz, &fc, &wc);
			}
			float ff[3];
			ff[0] = group_ff(gmask, part, 0, splittable), ff[1] = group_ff(gmask, part, 1, splittable);
			ff[2] = group_ff(gmask, part, 2, splittable);
			const float ffs = group_ff(gmask, part, 3, splittable);
			bool split = false;
			if (try_split) {
				const float ground0 = (float)((double)(ff[0] + ff[1] + ff[2]) / 3.0);
				float dlast = fabsf(ground0 - ffs);	// deltaff of the serial loop after its last evaluation
				split = true;
				if (dlast < S.delta_ff) {
					const int samples = (int)(area / S.sample_area);
					bool exceeded = false;
					for (int kb = 0; kb < samples && !exceeded; kb += kSplitBatch) {
						const int k = kb + q;
						const bool mine = q < kSplitBatch && k < samples;
						float ground = 0.0f, spart = 0.0f;
						FFCounters f1 = {0, 0, 0};
						WalkCounters w1 = {0, 0, 0, 0, 0, 0, 0};
						if (mine) {	// GetPointInTriangle (shade.cpp:49-68) with hashed weights
							float r1 = (float)(hash3((unsigned)S.shot_no, (unsigned)i, 3u * k) >> 1);
							float r2 = (float)(hash3((unsigned)S.shot_no, (unsigned)i, 3u * k + 1u) >> 1);
							float r3 = (float)(hash3((unsigned)S.shot_no, (unsigned)i, 3u * k + 2u) >> 1);
							const float tot = r1 + r2 + r3;
							r1 /= tot, r2 /= tot, r3 /= tot;
							const float sx = r1 * g0.x + r2 * g1.x + r3 * g2.x, sy = r1 * g0.y + r2 * g1.y + r3 * g2.y,
								    sz = r1 * g0.z + r2 * g1.z + r3 * g2.z;
							ground = r1 * ff[0] + r2 * ff[1] + r3 * ff[2];
							spart = adapt_ff_sub<true>(S.tree, sh, sub, pn.x, pn.y, pn.z, pn.w, parent, sx, sy, sz, &f1, &w1);
						}
						const float f = group_ff(gmask, spart, q < kSplitBatch ? q : 0, splittable);
						const float d = mine ? fabsf(ground - f) : 0.0f;
						const bool ex = mine && d > S.delta_ff;
						const unsigned exm = (__ballot_sync(gmask, ex) >> gshift) & ((1u << kSplitGroup) - 1u);
						const int first = exm ? (__ffs(exm) - 1) / 3 : kSplitBatch;	// sample at which the serial loop breaks
						if (mine && q <= first) add_counters(fc, wc, f1, w1);
						if (exm) {
							exceeded = true;
						} else {
							const int last = min(kSplitBatch - 1, samples - 1 - kb);
							dlast = __shfl_sync(gmask, d, 3 * last, kSplitGroup);
						}
					}
					if (!exceeded && dlast < S.delta_ff) split = false;	// smooth enough
				}
				if (split) {	// PartitionTriangleWithCenterAndStore's shape refusals (shade.cpp:265-275)
					float maxLen = 0.0f, minLen = 1e+30f;
					const float4 *g[3] = {&g0, &g1, &g2};
					for (int v = 0; v < 3; v++) {
						const float dx = g[(v + 1) % 3]->x - g[v]->x, dy = g[(v + 1) % 3]->y - g[v]->y, dz = g[(v + 1) % 3]->z - g[v]->z;
						const float l = dx * dx + dy * dy + dz * dz;
						maxLen = (maxLen < l) ? l : maxLen;
						minLen = (l < minLen) ? l : minLen;
					}
					if ((double)(minLen / maxLen) < 1e-1 || minLen < 1) split = false;
				}
			}
			if (gl == 0) {
				if (!split) {
					split_cal_radiosity(sh, S.fk[i], ff, a0, a1, a2);
					S.st[0][i] = a0, S.st[1][i] = a1, S.st[2][i] = a2;
				} else {
					S.count[i] = (parent == i) ? 3 : 2;
					S.ffs[i] = make_float4(ff[0], ff[1], ff[2], ffs);
				}
			}
		}
	}
	flush_counters(S.stats, fc, wc, shaded);
}

// ---- exclusive scan of count[lo, hi): per-CTA sums, then one CTA scans the sums ----------------------
constexpr int kScanBlock = 256;
constexpr int kReshade = -2;	// count[] mark: child replaced in place, to be re-shaded by split_reshade_kernel

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
	__shared__ int wsum[kScanBlock / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const int t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += t;
	}
	if (lane == 31) wsum[warp] = inc;
	__syncthreads();
	int base = 0, all = 0;
	for (int w = 0; w < kScanBlock / 32; w++) {
		if (w < warp) base += wsum[w];
		all += wsum[w];
	}
	__syncthreads();
	*total = all;
	return base + inc - v;
}

__global__ void __launch_bounds__(kScanBlock) split_scan_sums_kernel(SplitScene S, int lo, int hi) {
	const int i = lo + blockIdx.x * kScanBlock + threadIdx.x;
	int total;
	block_exclusive_scan(i < hi ? S.count[i] : 0, &total);
	if (threadIdx.x == 0) S.block_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock) split_scan_offsets_kernel(SplitScene S, int nblocks, int *growth) {
	__shared__ int carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < nblocks; base += kScanBlock) {
		const int b = base + threadIdx.x;
		int total;
		const int ex = block_exclusive_scan(b < nblocks ? S.block_sum[b] : 0, &total);
		if (b < nblocks) S.block_off[b] = carry + ex;
		__syncthreads();
		if (threadIdx.x == 0) carry += total;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		*growth = carry;
		// every request of this round is granted when even the last one finds room (isFullPool(4), shade.cpp:262):
		// the pool size is advanced here; otherwise the host works out the granted prefix and uploads it
		const int n = *S.trinum;
		if (n + carry + 4 < S.limit) *S.trinum = n + carry;
	}
}

// ---- stream-compacted append (center split, shade.cpp:261-306 + 213-241) ------------------------------
__device__ void split_write_child(const SplitScene &S, int dst, int src, const float p[3][3], const float4 &a0, const float4 &a1,
				  const float4 &a2, int parent, int nb0, int nb1, int nb2) {
	// CalCenter / CalArea of the child (vector.h:54-62): normal left un-normalised
	const float e0[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
	const float e1[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
	const float nx = e0[1] * e1[2] - e1[1] * e0[2], ny = e0[2] * e1[0] - e1[2] * e0[0], nz = e0[0] * e1[1] - e1[0] * e0[1];
	const float nlen = sqrtf(nx * nx + ny * ny + nz * nz);
	const float area = nlen / 2;
	S.p[0][dst] = make_float4(p[0][0], p[0][1], p[0][2], __int_as_float(parent));
	S.p[1][dst] = make_float4(p[1][0], p[1][1], p[1][2], area);
	S.p[2][dst] = make_float4(p[2][0], p[2][1], p[2][2], 0.0f);
	S.n[dst] = make_float4(nx, ny, nz, nlen);
	S.st[0][dst] = a0, S.st[1][dst] = a1, S.st[2][dst] = a2;
	S.fk[dst] = S.fk[src];
	if (dst != src) {	// static fields travel with the whole-record copy (*t = *s)
		const int *from = (const int *)&S.rec[src];
		int *to = (int *)&S.rec[dst];
		for (int w = 0; w < (int)(sizeof(radk_triangle) / 4); w++) to[w] = from[w];
	}
	S.rec[dst].neighbor[0] = nb0 >= 0 ? nb0 : S.rec[dst].neighbor[0];
	S.rec[dst].neighbor[1] = nb1 >= 0 ? nb1 : S.rec[dst].neighbor[1];
	S.rec[dst].neighbor[2] = nb2 >= 0 ? nb2 : S.rec[dst].neighbor[2];
}

__global__ void __launch_bounds__(kScanBlock) split_append_kernel(SplitScene S, const ShooterRec *__restrict__ shp, int lo, int hi, int base_trinum) {
	__shared__ ShooterRec sh;
	if (threadIdx.x < sizeof(ShooterRec) / 4) ((int *)&sh)[threadIdx.x] = ((const int *)shp)[threadIdx.x];
	__syncthreads();
	const int i = lo + blockIdx.x * kScanBlock + threadIdx.x;
	const int cnt = i < hi ? S.count[i] : 0;
	int total;
	const int off = S.block_off[blockIdx.x] + block_exclusive_scan(cnt, &total);
	FFCounters fc = {0, 0, 0};
	WalkCounters wc = {0, 0, 0, 0, 0, 0, 0};
	unsigned nsplit = 0u;
	if (cnt > 0) {
		const int first = base_trinum + off;	// where the serial reference's trinum stood at this slot's turn
		const float4 g0 = S.p[0][i], g1 = S.p[1][i], g2 = S.p[2][i];
		const int parent = __float_as_int(g0.w);
		float4 a0 = S.st[0][i], a1 = S.st[1][i], a2 = S.st[2][i];
		const float4 f = S.ffs[i];
		if (first + 4 >= S.limit) {
			// isFullPool(4) at this slot's turn (shade.cpp:262): refused -> CalRadiosity with the vertex form factors
			const float ff[3] = {f.x, f.y, f.z};
			split_cal_radiosity(sh, S.fk[i], ff, a0, a1, a2);
			S.st[0][i] = a0, S.st[1][i] = a1, S.st[2][i] = a2;
		} else {
			nsplit = 1u;
			const float c[3] = {(float)((double)(g0.x + g1.x + g2.x) / 3.0), (float)((double)(g0.y + g1.y + g2.y) / 3.0),
					    (float)((double)(g0.z + g1.z + g2.z) / 3.0)};
			const float P[3][3] = {{g0.x, g0.y, g0.z}, {g1.x, g1.y, g1.z}, {g2.x, g2.y, g2.z}};
			const int t1 = first, t2 = first + 1, t3 = (cnt == 3) ? first + 2 : i;
			float q[3][3];
			// t1: vertex 0 -> centre; t2: vertex 1 -> centre; t3: vertex 2 -> centre; children keep the old vertices' accB
			for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) q[v][k] = P[v][k];
			for (int k = 0; k < 3; k++) q[0][k] = c[k];
			split_write_child(S, t1, i, q, a0, a1, a2, parent, t3, -1, t2);
			for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) q[v][k] = P[v][k];
			for (int k = 0; k < 3; k++) q[1][k] = c[k];
			split_write_child(S, t2, i, q, a0, a1, a2, parent, t3, t1, -1);
			for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) q[v][k] = P[v][k];
			for (int k = 0; k < 3; k++) q[2][k] = c[k];
			if (cnt == 3) {
				// an untouched original: becomes logical, all three children are shaded in the next round
				split_write_child(S, t3, i, q, a0, a1, a2, parent, -1, t1, t2);
				S.p[0][i].w = __int_as_float(-1);
				S.rec[i].parent = -1;
			} else {
				// replaced in place by t3 and shaded at once with ff2 = {ff0, ff1, FF at the new vertex} (shade.cpp:509-520):
				// split_reshade_kernel does that with three lanes per slot; the slot is marked for it
				split_write_child(S, t3, i, q, a0, a1, a2, parent, -1, t1, t2);
				S.count[i] = kReshade;
			}
		}
	}
	flush_counters(S.stats, fc, wc, 0u);
	const unsigned s = __reduce_add_sync(0xffffffffu, nsplit);
	if ((threadIdx.x & 31) == 0 && s) atomicAdd(&S.stats->splits, (unsigned long long)s);
}

// The in-place children of this round (count == kReshade): form factor at the new vertex (the centroid, vertex 2 of the
// child), CalRadiosity with {ff0, ff1, that} (shade.cpp:509-520).  Four lanes per slot, three of them walk.
__global__ void __launch_bounds__(128) split_reshade_kernel(SplitScene S, const ShooterRec *__restrict__ shp, int lo, int hi) {
	__shared__ ShooterRec sh;
	if (threadIdx.x < sizeof(ShooterRec) / 4) ((int *)&sh)[threadIdx.x] = ((const int *)shp)[threadIdx.x];
	__syncthreads();
	const int sub = threadIdx.x & 3;
	const int i = lo + (blockIdx.x * 128 + threadIdx.x) / 4;
	const unsigned gmask = 0xfu << ((threadIdx.x & 31) & ~3);
	FFCounters fc = {0, 0, 0};
	WalkCounters wc = {0, 0, 0, 0, 0, 0, 0};
	if (i < hi && S.count[i] == kReshade) {
		const float4 g0 = S.p[0][i], g2 = S.p[2][i], pn = S.n[i];
		const int parent = __float_as_int(g0.w);
		float part = 0.0f;
		if (sub < 3) part = adapt_ff_sub<true>(S.tree, sh, sub, pn.x, pn.y, pn.z, pn.w, parent, g2.x, g2.y, g2.z, &fc, &wc);
		const float v0 = __shfl_sync(gmask, part, 0, 4), v1 = __shfl_sync(gmask, part, 1, 4), v2 = __shfl_sync(gmask, part, 2, 4);
		if (sub == 0) {
			const float4 f = S.ffs[i];
			const float ff2[3] = {f.x, f.y, sh.splittable ? v1 + v2 : v0};
			float4 a0 = S.st[0][i], a1 = S.st[1][i], a2 = S.st[2][i];
			split_cal_radiosity(sh, S.fk[i], ff2, a0, a1, a2);
			S.st[0][i] = a0, S.st[1][i] = a1, S.st[2][i] = a2;
		}
	}
	flush_counters(S.stats, fc, wc, 0u);
}

// end of a shot: the shooter's unshot energy is spent (rad.cpp:446-449), bookkeeping
__global__ void split_finish_shot_kernel(SplitScene S, const ShooterRec *shp) {
	if (!shp->valid) return;
	const int s = shp->slot;
	S.st[0][s].w = 0.0f, S.st[1][s].w = 0.0f, S.st[2][s].w = 0.0f;
	const long long k = (*S.sel.shots_done)++;
	S.sel.shot_log[k % S.sel.shot_log_cap] = s;
	atomicAdd(&S.stats->shots, 1ull);
}

}  // namespace radk
