// K3a, streaming version: batched vertex form factors with warp-autonomous work fetching (sm_100a).
//
// ff_batch_kernel (fast_kernels.cuh) gives every CTA one tile of 1024 vertices of one (row, half): about
// 500 rays for 256 lanes, i.e. two rays per lane, so lanes that drew short rays idle until the CTA's
// longest rays end (ncu: 18 of 32 lanes active in the leaf lookup, 10 in the triangle tests).  Here the
// grid is persistent and every WARP streams work on its own:
//
//   * work unit = (cache row j, shooter half, tile of kFSTile unique vertices), handed out by one global
//     counter in row-major order; a warp that cannot give every idle lane a ray grabs the next unit, runs
//     the facing culls + analytic form factor for it (4 vertices per lane) and appends the survivors to
//     its own ring in shared memory - lanes in the middle of a walk just keep their ray in registers;
//   * idle lanes take rays from the ring in order, so lanes stay busy until the whole launch runs dry and
//     the only tail is the last ray of the launch (no per-CTA and no per-wave tail);
//   * the leaf visit is leaf_visit() (rt_core.cuh): leaves that list the shooter's original and leaves
//     that do not run the same bounded any-hit scan, so lanes in the two kinds of leaves no longer
//     serialise (the exact nearest-hit scan took 9.6 % of the instructions at 5.9 active lanes).
//
// Same arithmetic and same results as ff_batch_kernel / vertex_ff_kernel (CalFF shade.cpp:107-127,
// RayHitted raycast.cpp:335-360); only the scheduling differs.
#pragma once
#include "fast_kernels.cuh"

namespace radk {

constexpr int kFSBlock = 256;
constexpr int kFSWarps = kFSBlock / 32;
constexpr int kFSTile = 64;	// vertices per work unit (2 per lane)
constexpr int kFSRing = 128;	// ring entries per warp: < 32 queued before a refill + at most kFSTile survivors
constexpr int kFSChunk = 16;	// units per chunk: a CTA's warps work through one chunk of kFSChunk * kFSTile neighbouring
				// vertices of one (row, half) together, so they share the leaf lists / occluders in L1
constexpr int kFSUBits = 26;	// ring entry = unique vertex | (j * 2 + half) << 26
constexpr unsigned kFSDone = 0xffffffu;

struct FSRow {	// what a lane needs to know about (row j of this launch, shooter half h)
	float cx, cy, cz;
	int logsrc;
	float nx, ny, nz, nlen;
	float area;
	int row;	// cache row
	int live;	// 0: second half of an unsplittable shooter (form factor 0 for every vertex)
	int pad;
};

// Next work unit of the calling warp (lane 0 only): units come from the CTA's current chunk; the warp that
// finds the chunk used up fetches the next chunk from the global counter.  *state = chunk << 8 | next unit.
// Returns chunk * kFSChunk + unit, or ~0u when the launch has no more chunks.
__device__ __forceinline__ unsigned grab_unit(unsigned *state, unsigned *global_next, unsigned nchunks) {
	for (;;) {
		const unsigned old = atomicAdd(state, 1u);
		const unsigned idx = old & 0xffu, ch = old >> 8;
		if (ch == kFSDone) return ~0u;
		if (idx < (unsigned)kFSChunk) return ch * kFSChunk + idx;
		if (idx == (unsigned)kFSChunk) {	// first to overrun: refill
			const unsigned g = atomicAdd(global_next, 1u);
			atomicExch(state, (g < nchunks ? g : kFSDone) << 8);
			continue;
		}
		while ((*((volatile unsigned *)state) >> 8) == ch) { }	// another warp is refilling
	}
}

template <int MINB, int UNR>
__global__ void __launch_bounds__(kFSBlock, MINB) ff_stream_kernel(DevScene S, FastState F) {
	__shared__ FSRow s_tab[2 * kMaxRows];
	__shared__ unsigned s_ring[kFSWarps][kFSRing];
	__shared__ unsigned s_state;
	const int nt = *F.ntodo;
	if (nt <= 0) return;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned tiles = (unsigned)((S.U + kFSTile - 1) / kFSTile);
	const unsigned cpr = (tiles + kFSChunk - 1) / kFSChunk;		// chunks per (row, half)
	const unsigned nchunks = (unsigned)nt * 2u * cpr;
	for (int k = threadIdx.x; k < 2 * nt; k += kFSBlock) {
		const int row = F.todo[k >> 1];
		const ShooterRec *r = F.rec + row;
		const FFSample sm = r->splittable ? r->half[k & 1] : r->centre;
		FSRow t;
		t.cx = sm.cx, t.cy = sm.cy, t.cz = sm.cz, t.logsrc = r->logsrc;
		t.nx = sm.nx, t.ny = sm.ny, t.nz = sm.nz, t.nlen = sm.nlen;
		t.area = sm.area, t.row = row, t.live = (r->splittable || (k & 1) == 0), t.pad = 0;
		s_tab[k] = t;
	}
	if (threadIdx.x == 0) {
		const unsigned g = atomicAdd(F.ff_work, 1u);
		s_state = (g < nchunks ? g : kFSDone) << 8;
	}
	__syncthreads();
	const DevTree &T = S.tree;
	const unsigned ltmask = (1u << lane) - 1u;
	unsigned *ring = s_ring[warp];
	unsigned head = 0, tail = 0;	// warp-uniform ring cursors
	bool more = true;		// units may remain
	// per-lane ray, kept small so that the culls of a refill fit beside it: the probe point is recomputed from t,
	// the form factor already sits in the cache row (phase 1 writes it, a blocked ray zeroes it)
	constexpr unsigned kNone = ~0u;
	unsigned ent = kNone;		// unique vertex | (j * 2 + half) << 26
	int ban = 0, prev = -1;
	unsigned sg = 0;		// leaf lookups so far | "precision exceeded" retries << 11
	float px = 0, py = 0, pz = 0, vx = 0, vy = 0, vz = 0, t = 0;
	for (;;) {
		const unsigned needm = __ballot_sync(0xffffffffu, ent == kNone);
		const int nneed = __popc(needm);
		while (more && (int)(tail - head) < nneed) {
			unsigned w = 0;
			if (lane == 0) w = grab_unit(&s_state, F.ff_work, nchunks);
			w = __shfl_sync(0xffffffffu, w, 0);
			if (w == ~0u) { more = false; break; }
			// chunk c = (j * cpr + chunk-in-row) * 2 + half: the two halves of the same vertices are neighbours in time
			const unsigned c = w / kFSChunk, k = w - c * kFSChunk;
			const unsigned jc = c >> 1, j = jc / cpr;
			const unsigned tile = (jc - j * cpr) * kFSChunk + k, jh = j * 2u + (c & 1u);
			const FSRow *tb = &s_tab[jh];
			float *out = (float *)(F.ff + (size_t)tb->row * S.U) + (jh & 1u);
			const int ubase = (int)(tile * kFSTile) + lane;
			if (!tb->live) {
#pragma unroll
				for (int it = 0; it < kFSTile / 32; it++)
					if (ubase + it * 32 < S.U) out[2 * (size_t)(ubase + it * 32)] = 0.0f;
				continue;
			}
			FFSample smp;
			smp.cx = tb->cx, smp.cy = tb->cy, smp.cz = tb->cz, smp.nx = tb->nx, smp.ny = tb->ny, smp.nz = tb->nz;
			smp.area = tb->area, smp.nlen = tb->nlen;
#pragma unroll(UNR)
			for (int it = 0; it < kFSTile / 32; it++) {
				const int u = ubase + it * 32;
				bool live = false;
				if (u < S.U) {
					const float4 v = __ldg(&S.uv[u]);
					const float4 pn = __ldg(&S.par_n[__float_as_int(v.w)]);
					float dx, dy, dz;
					const float ff = ff_geom(smp, pn.x, pn.y, pn.z, pn.w, v.x, v.y, v.z, dx, dy, dz, live);
					out[2 * (size_t)u] = live ? ff : 0.0f;	// the value if the ray reaches the shooter
				}
				const unsigned ball = __ballot_sync(0xffffffffu, live);
				if (live) ring[(tail + __popc(ball & ltmask)) & (kFSRing - 1)] = (unsigned)u | (jh << kFSUBits);
				tail += __popc(ball);
			}
			__syncwarp();
		}
		const int avail = (int)(tail - head);
		if (ent == kNone) {
			const int rank = __popc(needm & ltmask);
			if (rank < avail) {
				ent = ring[(head + rank) & (kFSRing - 1)];
				const float4 v = __ldg(&S.uv[ent & ((1u << kFSUBits) - 1u)]);
				const FSRow *tb = &s_tab[ent >> kFSUBits];
				px = v.x, py = v.y, pz = v.z, ban = __float_as_int(v.w);
				vx = tb->cx - px, vy = tb->cy - py, vz = tb->cz - pz;
				prev = -1, sg = 0, t = 0.f;
			}
		}
		head += (unsigned)(nneed < avail ? nneed : avail);
		__syncwarp();
		if (!__any_sync(0xffffffffu, ent != kNone)) break;
		// ---- one leaf step per lane (RayHitted, raycast.cpp:335-360), in phases that end with a warp-wide
		// reconvergence: lanes that leave a leaf list early wait for the scans of the others instead of running
		// the next phase on their own
		const int u = (int)(ent & ((1u << kFSUBits) - 1u));
		int verdict = -1;	// -1 undecided, 0 blocked / left the scene, 1 shooter visible
		int cur = -1;
		bool visit = false;
		if (ent != kNone) {	// phase A: leaf lookup
			if ((sg & 0x7ffu) == 0) cur = __ldg(&S.uv_leaf[u]);	// the first probe point is the vertex itself: its leaf is static
			else cur = leaf_of<false>(T, px + t * vx, py + t * vy, pz + t * vz, nullptr);
			sg++;
			if (cur < 0 || (sg & 0x7ffu) > (unsigned)kMaxWalkSteps) {
				verdict = 0;
			} else if (prev == cur) {		// raycast.cpp:344-351
				t = (float)((double)t + 1.0);
				sg += 1u << 11;
				if ((sg >> 11) > 100000u) verdict = 0;
			} else {
				visit = true;
			}
		}
		__syncwarp();
		if (visit) {	// phase B: the leaf's list
			const FSRow *tb = &s_tab[ent >> kFSUBits];
			const unsigned mw = __ldg(&F.row_mask[(size_t)tb->row * F.mask_words + (cur >> 5)]);
			verdict = leaf_visit(T, cur, (mw >> (cur & 31)) & 1u, px, py, pz, vx, vy, vz, ban, tb->logsrc);
		}
		__syncwarp();
		if (visit && verdict < 0) {	// phase C: NextPoint (raycast.cpp:235-251)
			const float4 lo = __ldg(&T.leaf_lo[cur]);
			const float4 hi = __ldg(&T.leaf_hi[cur]);
			prev = cur;
			float tn = 1000000.0f, s;
			if (fabsf(vx) > 1e-6f) { s = ((vx > 0 ? hi.x : lo.x) - px) / vx; tn = (s < tn) ? s : tn; }
			if (fabsf(vy) > 1e-6f) { s = ((vy > 0 ? hi.y : lo.y) - py) / vy; tn = (s < tn) ? s : tn; }
			if (fabsf(vz) > 1e-6f) { s = ((vz > 0 ? hi.z : lo.z) - pz) / vz; tn = (s < tn) ? s : tn; }
			t = (float)((double)tn + 0.1);
		}
		if (verdict >= 0) {	// phase D: the ray is decided; a visible ray keeps the form factor phase 1 stored
			if (verdict == 0) {
				const unsigned jh = ent >> kFSUBits;
				float *out = (float *)(F.ff + (size_t)s_tab[jh].row * S.U) + (jh & 1u);
				out[2 * (size_t)u] = 0.0f;
			}
			ent = kNone;
		}
	}
}

}  // namespace radk
