// Device core of the shooting loop: octree walk, triangle test, form factor.
//
// Arithmetic contract: every expression follows the reference's C++
// (src/raycast.cpp:235-360, src/shade.cpp:107-168) operation by operation - float unless the
// reference promotes through a double literal - evaluated left to right.  The translation unit
// is compiled with -fmad=false (no contraction), IEEE division and square root, no FTZ, so the
// results are bit-identical to the reference built with -ffp-contract=off.
//
// The functions are __host__ __device__ so that tests/hostemu can run exactly this code on the
// CPU against the oracle when no GPU is present; the product only ever calls them from kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define RADK_HD __host__ __device__ __forceinline__
#else
#define RADK_HD inline
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
#endif

namespace radk {

// ---- flattened scene, as laid out in HBM by scene_upload (ctx.cu) -------------------------------
// The reference's pointer-chasing root descent (TreeNodeNum, raycast.cpp:266-278) is replaced by a
// dense table over the finest grid: cell (gx,gy,gz) -> leaf node, built on the host by running the
// reference descent for every cell, so the leaf found is the same by construction.
struct DevTree {
	const int *cell_leaf;		// [(grid_num+1)^3], x fastest
	const float4 *leaf_box;		// 2 per node, one 32-byte sector: g0.xyz | list offset (int bits), g1.xyz | depth (int
					// bits, statistics only) - see leaf_lo() / leaf_hi()
	const int *list;		// leaf lists, -1 terminated, ascending triangle index
	const int4 *list4;		// the same lists, each padded with -1 to a multiple of 4 entries (16 B aligned)
	const int *leaf_list4;		// per node: offset (in int4 units) of its padded list, -1 for inner nodes
	const int *tri_leaf_begin;	// [n0+1] inverted index: leaves whose list contains triangle i ...
	const int *tri_leaves;		// ... are tri_leaves[tri_leaf_begin[i] .. tri_leaf_begin[i+1])
	int nnodes;
	const float4 *occ;		// plane parts of all occluders (2 float4 each), then their side-plane parts (3 float4
					// each): see occ_plane() / occ_side()
	int nocc;			// number of occluders (= n0)
	float G0[3], G1[3];
	float min_grid_len;
	int grid_num;
	double inv_min_grid_len;	// 1.0 / (double)min_grid_len, see cell_index
	float inv_min_grid_len_f;	// the same in float, and the guard band (in cells) within which the float
	float cell_eps;			// estimate of a cell coordinate is not trusted (>= 0.25: never trusted)
};
// occluder record (the {n, c, se} of TriStore[idx] at shot time, raycast.cpp:288-309):
//   plane part  q0 = n.x n.y n.z c.x | q1 = c.y c.z se00 se01
//   side part   q2 = se02 se03 se10 se11 | q3 = se12 se13 se20 se21 | q4 = se22 se23 - -
// Most tests end in the plane part (sign of t, t range), so the plane parts are stored densely (four per
// 128-byte line instead of 1.6 when the 80-byte records were contiguous): more of them stay in L1.
constexpr int OCC_PLANE = 2, OCC_SIDE = 3;
// A walk visits at most ~3 * grid_num leaves.  When GridNum is not a power of two (float rounding of
// (G1-G0)/MinGridLen, e.g. 63) the reference's cell -> leaf descent disagrees with the leaf boxes and its
// RayHitted can bounce between two leaves forever; the walk is cut after this many leaf lookups and the
// ray counts as leaving the scene (the oracle applies the same cut, the reference itself would hang).
constexpr int kMaxWalkSteps = 1024;

struct WalkCounters {
	unsigned traced, leaf_visits, descent_steps, tri_tests, tri_accepts, precision_exceeded, guard_trips;
};

RADK_HD int float_as_int_hd(float f) {
#if defined(__CUDA_ARCH__)
	return __float_as_int(f);
#else
	union { float f; int i; } u; u.f = f; return u.i;
#endif
}
RADK_HD float int_as_float_hd(int i) {
#if defined(__CUDA_ARCH__)
	return __int_as_float(i);
#else
	union { float f; int i; } u; u.i = i; return u.f;
#endif
}
template <typename T> RADK_HD T ldg(const T *p) {
#if defined(__CUDA_ARCH__)
	return __ldg(p);
#else
	return *p;
#endif
}

RADK_HD const float4 *occ_plane(const DevTree &T, int idx) { return T.occ + (size_t)idx * OCC_PLANE; }
RADK_HD const float4 *occ_side(const DevTree &T, int idx) { return T.occ + (size_t)T.nocc * OCC_PLANE + (size_t)idx * OCC_SIDE; }
RADK_HD float4 leaf_lo(const DevTree &T, int node) { return ldg(&T.leaf_box[2 * (size_t)node]); }
RADK_HD float4 leaf_hi(const DevTree &T, int node) { return ldg(&T.leaf_box[2 * (size_t)node + 1]); }

// vector.h:36 InnerProd
RADK_HD float dot3(float ax, float ay, float az, float bx, float by, float bz) {
	return ax * bx + ay * by + az * bz;
}

// gnum = (int)(((double)q - G0) / MinGridLen) (raycast.cpp:263).  The double division is replaced by a
// multiplication with the rounded reciprocal; both results differ from the exact quotient by less
// than 1e-13, so their truncations can only disagree when the quotient lies within 1e-9 of an
// integer - in that case the reference's division is evaluated instead.  A float estimate with its own
// (much wider) guard band is tried before any double-precision work.  Bit-identical by construction.
RADK_HD int cell_index(const DevTree &T, float q, float g0) {
	// float estimate first: its error is below cell_eps cells (three roundings of relative size 2^-24 on a
	// value of at most grid_num + 1), so away from cell boundaries its truncation is the reference's
	const float yf = (q - g0) * T.inv_min_grid_len_f;
	const int nf = (int)yf;
	const float ff = yf - (float)nf;
	if (T.cell_eps < 0.25f && ff > T.cell_eps && ff < 1.0f - T.cell_eps) return nf;
	const double x = (double)q - (double)g0;
	const double y = x * T.inv_min_grid_len;
	int n = (int)y;
	const double fr = y - (double)n;
	if (fr < 1e-9 || fr > 0.999999999) n = (int)(x / (double)T.min_grid_len);
	return n;
}

// TreeNodeNum (raycast.cpp:254-281): leaf containing q, or -1 outside the scene cube.
template <bool STATS>
RADK_HD int leaf_of(const DevTree &T, float qx, float qy, float qz, WalkCounters *wc) {
	if (qx < T.G0[0] || qx > T.G1[0] || qy < T.G0[1] || qy > T.G1[1] || qz < T.G0[2] || qz > T.G1[2])
		return -1;
	int gx = cell_index(T, qx, T.G0[0]);
	int gy = cell_index(T, qy, T.G0[1]);
	int gz = cell_index(T, qz, T.G0[2]);
	// every index >= grid_num takes the all-upper path of the descent, negatives the all-lower one
	const int g = T.grid_num, dim = g + 1;
	gx = gx < 0 ? 0 : (gx > g ? g : gx);
	gy = gy < 0 ? 0 : (gy > g ? g : gy);
	gz = gz < 0 ? 0 : (gz > g ? g : gz);
	const int leaf = ldg(&T.cell_leaf[(gz * dim + gy) * dim + gx]);
	if (STATS) wc->descent_steps += (unsigned)float_as_int_hd(leaf_hi(T, leaf).w);
	return leaf;
}

// TriHitted (raycast.cpp:283-313)
RADK_HD bool tri_hit(const float4 *o, const float4 *side, float px, float py, float pz, float vx, float vy, float vz, float &t_min) {
	const float4 q0 = ldg(o), q1 = ldg(o + 1);
	const float a1 = dot3(vx, vy, vz, q0.x, q0.y, q0.z);
	const float a2 = dot3(q0.w - px, q1.x - py, q1.y - pz, q0.x, q0.y, q0.z);
	// t = a2/a1 < 0.0001 is certain when the signs differ (or a2 is zero): skip the division.
	if ((a1 > 0.0f && a2 <= 0.0f) || (a1 < 0.0f && a2 >= 0.0f)) return false;
	const float t = a2 / a1;
	if (t <= 0.0001f) return false;	// == ((double)t < 0.0001): 0.0001f is the largest float below the double 0.0001
	if (t >= t_min) return false;
	const float hx = px + t * vx, hy = py + t * vy, hz = pz + t * vz;
	const float4 q2 = ldg(side);
	if ((hx * q1.z + hy * q1.w + hz * q2.x + q2.y) < 0.0f) return false;
	const float4 q3 = ldg(side + 1);
	if ((hx * q2.z + hy * q2.w + hz * q3.x + q3.y) < 0.0f) return false;
	const float4 q4 = ldg(side + 2);
	if ((hx * q3.z + hy * q3.w + hz * q4.x + q4.y) < 0.0f) return false;
	t_min = t;
	return true;
}

// RayHitted (raycast.cpp:335-360) with TriListHitted (316-333) and NextPoint (235-251) inlined.
// Returns the triangle index hit in the first leaf (in stepping order) that has any hit, or -1.
template <bool STATS>
RADK_HD int ray_walk(const DevTree &T, float px, float py, float pz, float vx, float vy, float vz, int ban,
		     WalkCounters *wc) {
	float t = 0.f;
	int prev = -1, cur;
	float qx = px, qy = py, qz = pz;
	int guard = 0, steps = 0;
	if (STATS) wc->traced++;
	while ((cur = leaf_of<STATS>(T, qx, qy, qz, wc)) >= 0) {
		if (++steps > kMaxWalkSteps) {
			if (STATS) wc->guard_trips++;
			return -1;
		}
		if (prev == cur) {		// raycast.cpp:344-351 "Precision Exceeded"
			if (STATS) wc->precision_exceeded++;
			t = (float)((double)t + 1.0);
			qx = px + t * vx, qy = py + t * vy, qz = pz + t * vz;
			if (++guard > 100000) {	// the reference would spin forever
				if (STATS) wc->guard_trips++;
				return -1;
			}
			continue;
		}
		const float4 lo = leaf_lo(T, cur);
		const float4 hi = leaf_hi(T, cur);
		if (STATS) wc->leaf_visits++;
		// TriListHitted
		int hit = -1;
		float t_min = 1000000.0f;
		const int *lp = T.list + float_as_int_hd(lo.w);
		int idx;
		while ((idx = ldg(lp++)) >= 0) {
			if (idx == ban) continue;
			if (STATS) wc->tri_tests++;
			if (tri_hit(occ_plane(T, idx), occ_side(T, idx), px, py, pz, vx, vy, vz, t_min)) {
				if (STATS) wc->tri_accepts++;
				hit = idx;
			}
		}
		if (hit >= 0) return hit;
		prev = cur;
		// NextPoint; fabsf(v) > 1e-6f == (fabs((double)v) > 1e-6): 1e-6f is below the double 1e-6
		float tn = 1000000.0f, s;
		if (fabsf(vx) > 1e-6f) { s = ((vx > 0 ? hi.x : lo.x) - px) / vx; tn = (s < tn) ? s : tn; }
		if (fabsf(vy) > 1e-6f) { s = ((vy > 0 ? hi.y : lo.y) - py) / vy; tn = (s < tn) ? s : tn; }
		if (fabsf(vz) > 1e-6f) { s = ((vz > 0 ? hi.z : lo.z) - pz) / vz; tn = (s < tn) ? s : tn; }
		tn = (float)((double)tn + 0.1);
		qx = px + tn * vx, qy = py + tn * vy, qz = pz + tn * vz;
		t = tn;
	}
	return -1;
}

// Is the shooter visible?  Same answer as (ray_walk(...) == logsrc), organised for the GPU:
//  * `has_src` is a bitmask over octree nodes: bit set <=> the leaf's list contains logsrc.  In a leaf
//    WITHOUT logsrc any accepted triangle ends the walk with "not visible" - which triangle is nearest
//    does not matter, and whether *some* triangle is accepted does not depend on the scan order
//    (the first acceptance always happens against t_min = 1e6).  Those leaves are scanned four list
//    entries at a time with independent tests (no loop-carried t_min), stopping at the first accept.
//  * a leaf WITH logsrc runs the reference's sequential nearest-hit scan unchanged.
// list4 entries carry a flag for occluders whose three side planes are all zero (the re-split
// originals of PrePartitionTriangles, SURVEY.md Fact 4): their inside test always passes, so the
// plane part of TriHitted decides alone and the side-plane loads are skipped.
constexpr int kPlaneFlag = 1 << 30;
constexpr int kIdMask = kPlaneFlag - 1;

// ---- bounded any-hit: one code path for leaves with and without the shooter's original ----------------
// TriListHitted (raycast.cpp:316-333) returns the accepted entry with the smallest t, the earliest list
// position (= smallest index, the lists are ascending) winning ties, because a later entry with
// t >= t_min is rejected.  With Acc = {entries that pass t > 0.0001, t < 1e6 and the inside test} the
// shooter's original `logsrc` is returned iff logsrc is in Acc and no other e in Acc has
// t_e < t_src or (t_e == t_src and e < logsrc).  So a leaf that lists logsrc is handled as: test logsrc
// alone (against 1e6), then scan the list for ANY entry accepted under the bound (t_max, tie) =
// (t_src, logsrc); a leaf that does not list it is the same scan with the bound (1e6, -1), which is
// the plain "t >= 1e6 -> reject".  Both kinds of leaves therefore run the same instructions and the
// scan stops at the first blocker.  The reference's behaviour once a NaN t (0/0) has been accepted is
// order dependent (every later comparison with t_min is false); such a visit returns 2 and the caller
// falls back to the sequential scan.
RADK_HD int any_hit_bounded_one(const DevTree &T, int entry, float px, float py, float pz, float vx, float vy, float vz,
				int ban, float t_max, int tie) {
	const int idx = entry & kIdMask;
	if (idx == ban) return 0;
	const float4 *o = occ_plane(T, idx);
	const float4 q0 = ldg(o), q1 = ldg(o + 1);
	const float a1 = dot3(vx, vy, vz, q0.x, q0.y, q0.z);
	const float a2 = dot3(q0.w - px, q1.x - py, q1.y - pz, q0.x, q0.y, q0.z);
	if ((a1 > 0.0f && a2 <= 0.0f) || (a1 < 0.0f && a2 >= 0.0f)) return 0;	// t <= 0: certain rejection
	const float t = a2 / a1;
	if (t <= 0.0001f) return 0;
	if (t > t_max || (t == t_max && idx >= tie)) return 0;
	if (!(entry & kPlaneFlag)) {	// zero side planes: the inside test cannot fail
		const float hx = px + t * vx, hy = py + t * vy, hz = pz + t * vz;
		const float4 *sd = occ_side(T, idx);
		const float4 q2 = ldg(sd);
		if ((hx * q1.z + hy * q1.w + hz * q2.x + q2.y) < 0.0f) return 0;
		const float4 q3 = ldg(sd + 1);
		if ((hx * q2.z + hy * q2.w + hz * q3.x + q3.y) < 0.0f) return 0;
		const float4 q4 = ldg(sd + 2);
		if ((hx * q3.z + hy * q3.w + hz * q4.x + q4.y) < 0.0f) return 0;
	}
	return t != t ? 2 : 1;
}

// Can this entry give t = 0 / 0 ?  (a1 == 0 and a2 == 0: the ray starts in the occluder's plane and runs inside it.)
// Such a t passes every test of TriHitted, also the inside test (all comparisons with NaN are false).
RADK_HD bool entry_gives_nan(const DevTree &T, int entry, float px, float py, float pz, float vx, float vy, float vz, int ban) {
	const int idx = entry & kIdMask;
	if (idx == ban) return false;
	const float4 *o = occ_plane(T, idx);
	const float4 q0 = ldg(o), q1 = ldg(o + 1);
	const float a2 = dot3(q0.w - px, q1.x - py, q1.y - pz, q0.x, q0.y, q0.z);
	if (a2 != 0.0f && a2 == a2) return false;	// almost always: the origin is not in the plane
	const float a1 = dot3(vx, vy, vz, q0.x, q0.y, q0.z);
	const float t = a2 / a1;
	return t != t;
}

// Does any entry of the padded list with index below `below` give a NaN t for this ray?
RADK_HD bool list_gives_nan_below(const DevTree &T, int list4_off, int below, float px, float py, float pz, float vx, float vy, float vz,
				  int ban) {
	const int4 *l4 = T.list4 + list4_off;
	for (;;) {
		const int4 id = ldg(l4++);
		const int e[4] = {id.x, id.y, id.z, id.w};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (int k = 0; k < 4; k++) {
			if (e[k] < 0) return false;
			if ((e[k] & kIdMask) < below && entry_gives_nan(T, e[k], px, py, pz, vx, vy, vz, ban)) return true;
		}
	}
}

// The exact sequential scan (TriListHitted, raycast.cpp:316-333) compared with logsrc; only reached when a
// NaN t was accepted in a leaf that lists logsrc, so it is kept out of line (registers of the hot path).
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
int leaf_visit_exact(const int *lp, const float4 *occ, int nocc, float px, float py, float pz, float vx, float vy, float vz, int ban, int logsrc) {
	int hit = -1;
	float t_min = 1000000.0f;
	int idx;
	while ((idx = ldg(lp++)) >= 0) {
		if (idx == ban) continue;
		if (tri_hit(occ + (size_t)idx * OCC_PLANE, occ + (size_t)nocc * OCC_PLANE + (size_t)idx * OCC_SIDE, px, py, pz, vx, vy, vz, t_min)) hit = idx;
	}
	return hit < 0 ? -1 : (hit == logsrc);
}

// leaf_visit() met a blocker (nan_idx < 0) or a NaN entry (nan_idx = its index) in a leaf whose logsrc passes: 0 unless a
// NaN entry below logsrc may have revived it - then the reference's sequential scan decides.  Out of line: registers of
// the hot scan.
RADK_HD int leaf_visit_rare(const DevTree &T, int cur, int nan_idx, float px, float py, float pz, float vx, float vy, float vz, int ban, int logsrc) {
	if (nan_idx >= 0 ? nan_idx > logsrc : !list_gives_nan_below(T, ldg(&T.leaf_list4[cur]), logsrc, px, py, pz, vx, vy, vz, ban)) return 0;
	return leaf_visit_exact(T.list + float_as_int_hd(leaf_lo(T, cur).w), T.occ, T.nocc, px, py, pz, vx, vy, vz, ban, logsrc);
}

// One leaf visit of a shooter-visibility ray: -1 nothing hit here (walk on), 0 the walk ends on another
// triangle, 1 the walk ends on logsrc.  Same answer as TriListHitted(...) compared with logsrc.
//
// NaN: an entry whose plane contains the ray (a1 == a2 == 0, t = 0/0) passes every test of TriHitted, leaves t_min = NaN
// and with it every later passing entry is accepted whatever its t (raycast.cpp:300,323-328); the lists are ascending,
// so "later" = larger triangle index.  Church has many such rays (vertices on a crease, shot along the adjoining wall).
// What the scan below relies on, with `passes` = t >= 0.0001 and the inside test, no bound on t:
//   * a NaN entry is always accepted: a leaf with one ends the walk;
//   * the walk can only end on logsrc if logsrc is listed, is not the banned triangle and passes, and then only if
//     no NaN entry has a larger index;
// so a NaN entry or a blocker met while logsrc does not pass means "ends elsewhere" without looking further, a NaN
// entry above logsrc too, and a blocker stands unless some NaN entry lies below logsrc.  Whatever is left (a NaN entry
// below a passing logsrc) takes the reference's sequential scan.
RADK_HD int leaf_visit(const DevTree &T, int cur, bool lists_src, float px, float py, float pz, float vx, float vy, float vz,
		       int ban, int logsrc, bool exact = false) {
	// exact: take the reference's sequential scan unconditionally (tests: the two scans must agree on every ray)
	if (exact) return leaf_visit_exact(T.list + float_as_int_hd(leaf_lo(T, cur).w), T.occ, T.nocc, px, py, pz, vx, vy, vz, ban, logsrc);
	// (t_max, tie): the bound of the any-hit scan.  tie >= 0 <=> logsrc is listed, not banned, passes and has t < 1e6.
	float t_max = int_as_float_hd(0x7fc00000);	// NaN: no bound at all for the source itself (`passes` above; even t = +inf passes once t_min is NaN)
	int tie = -1;
	if (lists_src && logsrc != ban && tri_hit(occ_plane(T, logsrc), occ_side(T, logsrc), px, py, pz, vx, vy, vz, t_max)) {
		if (!(t_max < 1000000.0f))	// t NaN, infinite or beyond the reference's initial t_min = 1e6 (rays inside the source's plane): sequential scan
			return leaf_visit_exact(T.list + float_as_int_hd(leaf_lo(T, cur).w), T.occ, T.nocc, px, py, pz, vx, vy, vz, ban, logsrc);
		tie = logsrc;
	} else {
		t_max = 1000000.0f;
	}
	const int4 *l4 = T.list4 + ldg(&T.leaf_list4[cur]);
	for (;;) {
		const int4 id = ldg(l4++);
		const int e[4] = {id.x, id.y, id.z, id.w};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (int k = 0; k < 4; k++) {
			if (e[k] < 0) return tie >= 0 ? 1 : -1;
			const int r = any_hit_bounded_one(T, e[k], px, py, pz, vx, vy, vz, ban, t_max, tie);
			if (r == 0) continue;
			if (tie < 0) return 0;	// accepted (blocker or NaN entry) while logsrc does not pass: the walk ends here, and not on logsrc
			return leaf_visit_rare(T, cur, r == 2 ? (e[k] & kIdMask) : -1, px, py, pz, vx, vy, vz, ban, logsrc);
		}
	}
}

// Is the shooter visible?  Same answer as (ray_walk(...) == logsrc); the walk of ff_batch_kernel.
template <typename MaskPtr>
RADK_HD bool ray_sees(const DevTree &T, MaskPtr has_src, float px, float py, float pz, float vx, float vy, float vz,
			      int ban, int logsrc, bool exact = false) {
	float t = 0.f;
	int prev = -1, cur;
	float qx = px, qy = py, qz = pz;
	int guard = 0, steps = 0;
	while ((cur = leaf_of<false>(T, qx, qy, qz, nullptr)) >= 0) {
		if (++steps > kMaxWalkSteps) return false;
		if (prev == cur) {		// raycast.cpp:344-351
			t = (float)((double)t + 1.0);
			qx = px + t * vx, qy = py + t * vy, qz = pz + t * vz;
			if (++guard > 100000) return false;
			continue;
		}
		const int v = leaf_visit(T, cur, (has_src[cur >> 5] >> (cur & 31)) & 1u, px, py, pz, vx, vy, vz, ban, logsrc, exact);
		if (v >= 0) return v == 1;
		const float4 lo = leaf_lo(T, cur);
		const float4 hi = leaf_hi(T, cur);
		prev = cur;
		float tn = 1000000.0f, s;
		if (fabsf(vx) > 1e-6f) { s = ((vx > 0 ? hi.x : lo.x) - px) / vx; tn = (s < tn) ? s : tn; }
		if (fabsf(vy) > 1e-6f) { s = ((vy > 0 ? hi.y : lo.y) - py) / vy; tn = (s < tn) ? s : tn; }
		if (fabsf(vz) > 1e-6f) { s = ((vz > 0 ? hi.z : lo.z) - pz) / vz; tn = (s < tn) ? s : tn; }
		tn = (float)((double)tn + 0.1);
		qx = px + tn * vx, qy = py + tn * vy, qz = pz + tn * vz;
		t = tn;
	}
	return false;
}

// ---- form factor ---------------------------------------------------------------------------------
// One shooter sample point (the fields CalFF reads from its source triangle, shade.cpp:107-127).
struct FFSample {
	float cx, cy, cz;
	float nx, ny, nz;
	float area;
	float nlen;	// norm(n): per-shot invariant, hoisted out of CalFF (same value, computed once)
};

struct FFCounters { unsigned calff, rays, visible; };

// The part of CalFF before the ray (shade.cpp:109-123): facing culls and the analytic point-to-patch
// form factor.  Returns 0 when the sample is culled, otherwise the (positive, possibly NaN) form
// factor that CalFF returns if the ray then reaches the shooter; dir = sample centre - p.
RADK_HD float ff_geom(const FFSample &S, float ndx, float ndy, float ndz, float desLen, float px, float py, float pz,
		      float &dx, float &dy, float &dz, bool &cast) {
	dx = S.cx - px, dy = S.cy - py, dz = S.cz - pz;
	cast = false;
	const float dotSrc = dot3(dx, dy, dz, S.nx, S.ny, S.nz);
	const float dotDes = -dot3(dx, dy, dz, ndx, ndy, ndz);
	if ((dotSrc < 0) != (dotDes < 0)) return 0.0f;
	const float n2 = dx * dx + dy * dy + dz * dz;
	const float dirLen = sqrtf(n2);
	const float ct1 = dotSrc / dirLen / S.nlen;
	const float ct2 = dotDes / dirLen / desLen;
	float ff = ct1 * ct2;
	if (ff <= 0.0f) return 0.0f;
	ff = (float)((double)ff * ((double)S.area / ((double)n2 * 3.1415926535897932)));
	if (ff <= 0.0f) return 0.0f;
	cast = true;	// NaN form factors pass both tests and do cast a ray, like the reference
	return ff;
}

// CalFF (shade.cpp:107-127).  desLen = norm(receiver normal) is passed in (pure function of n).
template <bool STATS>
RADK_HD float cal_ff(const DevTree &T, const FFSample &S, int logsrc, float ndx, float ndy, float ndz, float desLen,
		     int logdest, float px, float py, float pz, FFCounters *fc, WalkCounters *wc) {
	if (STATS) fc->calff++;
	float dx, dy, dz;
	bool cast;
	const float ff = ff_geom(S, ndx, ndy, ndz, desLen, px, py, pz, dx, dy, dz, cast);
	if (!cast) return 0.0f;
	if (STATS) fc->rays++;
	if (ray_walk<STATS>(T, px, py, pz, dx, dy, dz, logdest, wc) == logsrc) {
		if (STATS) fc->visible++;
		return ff;
	}
	return 0.0f;
}

// What the loop needs to know about the current shooter; written once per shot by the selection
// kernel (K1 + K2: listCanditate rad.cpp:387-404, AdaptCalFF's split shade.cpp:146-161,
// PartitionSource shade.cpp:76-99) and read by every receiver thread.
struct ShooterRec {
	FFSample centre;	// the triangle itself (the sample whose result AdaptCalFF discards)
	FFSample half[2];	// t1 = {m, p[e+2], p[e]}, t2 = {m, p[e+1], p[e+2]}
	float dB[3];		// unshot radiosity being shot
	float delta;		// (dB0+dB1)+dB2
	int splittable;		// longest edge^2 >= 0.1 (shade.cpp:154)
	int logsrc;		// original parent of the shooter = the index a visible ray must return
	int slot;		// TriStore slot (global)
	int recv;		// receiver index on the owning shard (-1 when owned by another rank)
	int valid;		// 0: no candidate (converged)
	int pad[3];
};

// K2: the shooter record of a receiver with vertices p, original normal n (|n| = nlen), area and
// unshot energy dB (AdaptCalFF's longest-edge choice shade.cpp:146-156, PartitionSource 76-99).
RADK_HD void make_shooter(const float p[3][3], float nx, float ny, float nz, float nlen, float area,
			  float d0, float d1, float d2, int parent, int slot, int recv, ShooterRec *out) {
	ShooterRec s;
	s.centre.cx = (float)((double)(p[0][0] + p[1][0] + p[2][0]) / 3.0);	// CalCenter, vector.h:54-57
	s.centre.cy = (float)((double)(p[0][1] + p[1][1] + p[2][1]) / 3.0);
	s.centre.cz = (float)((double)(p[0][2] + p[1][2] + p[2][2]) / 3.0);
	s.centre.nx = nx, s.centre.ny = ny, s.centre.nz = nz, s.centre.nlen = nlen;
	s.centre.area = area;
	int e = 0;
	float maxlen = 0.0f;
	for (int v = 0; v < 3; v++) {	// first maximum wins
		const float *a = p[v], *b = p[(v + 1) % 3];
		const float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
		const float l = dx * dx + dy * dy + dz * dz;
		if (maxlen < l) maxlen = l, e = v;
	}
	s.splittable = !((double)maxlen < 0.1);
	const int i0 = e, i1 = (e + 1) % 3, i2 = (e + 2) % 3;
	float m[3], c1[3], c2[3];
	for (int k = 0; k < 3; k++) m[k] = (float)((double)(p[i0][k] + p[i1][k]) / 2.0);
	for (int k = 0; k < 3; k++) {	// t1 = {m, p[i2], p[i0]}, t2 = {m, p[i1], p[i2]}
		c1[k] = (float)((double)(m[k] + p[i2][k] + p[i0][k]) / 3.0);
		c2[k] = (float)((double)(m[k] + p[i1][k] + p[i2][k]) / 3.0);
	}
	s.half[0] = s.centre, s.half[1] = s.centre;
	s.half[0].cx = c1[0], s.half[0].cy = c1[1], s.half[0].cz = c1[2];
	s.half[1].cx = c2[0], s.half[1].cy = c2[1], s.half[1].cz = c2[2];
	s.half[0].area = s.half[1].area = (float)((double)area / 2.0);
	s.dB[0] = d0, s.dB[1] = d1, s.dB[2] = d2;
	s.delta = d0 + d1 + d2;
	s.logsrc = parent;
	s.slot = slot;
	s.recv = recv;
	s.valid = 1;
	s.pad[0] = s.pad[1] = s.pad[2] = 0;
	*out = s;
}

// AdaptCalFF(CalFF(centre)) as it executes (shade.cpp:141-168, stray ';' at 164): for a splittable
// shooter the value is CalFF(t1)+CalFF(t2); the centre sample only matters for the statistics.
// adapt_ff_sub evaluates one of its three independent parts so that cooperating lanes can take one each:
//   sub 0: the centre sample (the value for an unsplittable shooter; statistics only otherwise),
//   sub 1 / 2: CalFF toward half t1 / t2 (0 for an unsplittable shooter).
template <bool STATS>
RADK_HD float adapt_ff_sub(const DevTree &T, const ShooterRec &sh, int sub, float ndx, float ndy, float ndz, float desLen,
			   int logdest, float px, float py, float pz, FFCounters *fc, WalkCounters *wc) {
	if (!sh.splittable)
		return sub == 0 ? cal_ff<STATS>(T, sh.centre, sh.logsrc, ndx, ndy, ndz, desLen, logdest, px, py, pz, fc, wc) : 0.0f;
	if (sub != 0) return cal_ff<STATS>(T, sh.half[sub - 1], sh.logsrc, ndx, ndy, ndz, desLen, logdest, px, py, pz, fc, wc);
	if (STATS) {	// reference-equivalent counting of the discarded centre sample (no walk counters)
		fc->calff++;
		const float dx = sh.centre.cx - px, dy = sh.centre.cy - py, dz = sh.centre.cz - pz;
		const float dotSrc = dot3(dx, dy, dz, sh.centre.nx, sh.centre.ny, sh.centre.nz);
		const float dotDes = -dot3(dx, dy, dz, ndx, ndy, ndz);
		if ((dotSrc < 0) == (dotDes < 0)) {
			const float n2 = dx * dx + dy * dy + dz * dz;
			const float dirLen = sqrtf(n2);
			float ff = (dotSrc / dirLen / sh.centre.nlen) * (dotDes / dirLen / desLen);
			if (!(ff <= 0.0f)) {
				ff = (float)((double)ff * ((double)sh.centre.area / ((double)n2 * 3.1415926535897932)));
				if (!(ff <= 0.0f)) {
					fc->rays++;
					// its visibility is the same walk the reference does; trace it only for the count
					if (ray_walk<false>(T, px, py, pz, dx, dy, dz, logdest, nullptr) == sh.logsrc) fc->visible++;
				}
			}
		}
	}
	return 0.0f;
}

template <bool STATS>
RADK_HD float adapt_ff(const DevTree &T, const ShooterRec &sh, float ndx, float ndy, float ndz, float desLen,
		       int logdest, float px, float py, float pz, FFCounters *fc, WalkCounters *wc) {
	const float f0 = adapt_ff_sub<STATS>(T, sh, 0, ndx, ndy, ndz, desLen, logdest, px, py, pz, fc, wc);
	if (!sh.splittable) return f0;
	const float f1 = adapt_ff_sub<STATS>(T, sh, 1, ndx, ndy, ndz, desLen, logdest, px, py, pz, fc, wc);
	const float f2 = adapt_ff_sub<STATS>(T, sh, 2, ndx, ndy, ndz, desLen, logdest, px, py, pz, fc, wc);
	return f1 + f2;
}

}  // namespace radk
