// Init-time mesh refinement: restates PrePartitionTriangles (reference src/shade.cpp:556-598),
// PartitionTriangleAndStore (shade.cpp:311-431) and the 2-way PartitionDestination
// (shade.cpp:171-211) AS THEY EXECUTE, because the result defines both the
// frozen receiver mesh and - through in-place overwrites of slots < n0 - the occluder records:
//  * largest-area-first longest-edge bisection, co-splitting the neighbour across that edge;
//  * the work set keeps stale (area, slot) entries of co-split neighbours; when such an entry is
//    popped the slot no longer is "an untouched original", so it is overwritten by its second child;
//  * children inherit n, Frgb, deltaB and interpolated accB; Brgb, se, deltaaccB keep the slot's
//    previous contents (zero for never-used slots);
//  * the cross-link test at shade.cpp:416 compares array addresses and is therefore always false.
#include <algorithm>
#include <cstdint>
#if defined(__linux__)
#include <sys/mman.h>
#endif
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iterator>
#include <map>
#include <set>
#include <vector>
#include <utility>
#include "model.hpp"

namespace radk {
namespace {

struct Splitter {
	Model &m;
	radk_triangle *T;	// the refinement pool (fixed for the whole run)

	bool pool_full(int need) const { return m.trinum + need >= m.pool_limit(); }	// shade.cpp:26-30
	int alloc() { return m.trinum++; }	// shade.cpp:14-25 (exhaustion is excluded by pool_full(4))

	void relink(int nb, int from, int to) {	// shade.cpp:33-43
		if (nb < 0) return;
		for (int v = 0; v < 3; v++)
			if (T[nb].neighbor[v] == from) T[nb].neighbor[v] = to;
	}

	static int longest_edge(const radk_triangle &t, float &len2) {
		int e = 0;
		len2 = 0.0f;
		for (int v = 0; v < 3; v++) {
			const float *a = t.p[v], *b = t.p[(v + 1) % 3];
			float d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
			float l = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
			if (len2 < l) len2 = l, e = v;
		}
		return e;
	}

	// shade.cpp:171-211: bisect s on edge e into a = {m, p[e+2], p[e]} and b = {m, p[e+1], p[e+2]}
	static void bisect(const radk_triangle &s, radk_triangle &a, radk_triangle &b, int e) {
		const int i0 = e, i1 = (e + 1) % 3, i2 = (e + 2) % 3;
		float mid[3], macc[3];
		for (int k = 0; k < 3; k++) {
			mid[k] = (float)((double)(s.p[i0][k] + s.p[i1][k]) / 2.0);
			macc[k] = (float)((double)(s.accB[i0][k] + s.accB[i1][k]) / 2.0);
		}
		for (int k = 0; k < 3; k++) {
			a.p[0][k] = mid[k], a.p[1][k] = s.p[i2][k], a.p[2][k] = s.p[i0][k];
			b.p[0][k] = mid[k], b.p[1][k] = s.p[i1][k], b.p[2][k] = s.p[i2][k];
			a.n[k] = b.n[k] = s.n[k];
			a.Frgb[k] = b.Frgb[k] = s.Frgb[k];
			a.deltaB[k] = b.deltaB[k] = s.deltaB[k];
			a.accB[0][k] = macc[k], a.accB[1][k] = s.accB[i2][k], a.accB[2][k] = s.accB[i0][k];
			b.accB[0][k] = macc[k], b.accB[1][k] = s.accB[i1][k], b.accB[2][k] = s.accB[i2][k];
		}
		for (int k = 0; k < 3; k++) {
			a.c[k] = (float)((double)(a.p[0][k] + a.p[1][k] + a.p[2][k]) / 3.0);
			b.c[k] = (float)((double)(b.p[0][k] + b.p[1][k] + b.p[2][k]) / 3.0);
		}
		a.area = b.area = (float)((double)s.area / 2.0);
		a.parent = b.parent = s.parent;
	}

	// Returns -1 refused, 0 split (an original became logical), 1 split (slot replaced in place).
	// out[] receives the slots now holding the (up to four) children, -1 where there is none.
	int split(int real, int out[4]) {
		out[0] = out[1] = out[2] = out[3] = -1;
		if (pool_full(4) || (double)T[real].area < 1e-8) return -1;
		float len2;
		const int e = longest_edge(T[real], len2);
		if (len2 < 1) return -1;
		const int nb = T[real].neighbor[e];
		if (nb >= 0 && (double)T[nb].area < 1e-1) return -1;

		int c1 = alloc(), c2 = alloc();
		bisect(T[real], T[c1], T[c2], e);
		const int outer1 = T[real].neighbor[(e + 2) % 3], outer2 = T[real].neighbor[(e + 1) % 3];
		T[c1].neighbor[1] = outer1;
		T[c2].neighbor[1] = outer2;
		T[c1].neighbor[2] = T[c2].neighbor[0] = -1;
		int replaced = 0;
		if (T[real].parent == real) {
			T[real].parent = -1;		// untouched original: becomes logical, geometry kept
		} else {
			T[real] = T[c2];		// whole-record copy: se/Brgb become the child's (zero/stale)
			c2 = real;
			m.trinum--;
			replaced = 1;
		}
		relink(outer1, real, c1);
		relink(outer2, real, c2);
		T[c1].neighbor[0] = c2;
		T[c2].neighbor[2] = c1;
		out[0] = c1, out[1] = c2;
		if (nb < 0) return replaced;

		int c3 = alloc(), c4 = alloc();
		int ne = 0;
		for (int v = 0; v < 3; v++)
			if (T[nb].neighbor[v] == real) ne = v;
		if (nb < real)				// "un-shade" (shade.cpp:391-396); deltaaccB is zero at init time
			for (int v = 0; v < 3; v++)
				for (int k = 0; k < 3; k++) T[nb].accB[v][k] -= T[nb].deltaaccB[v][k];
		bisect(T[nb], T[c3], T[c4], ne);
		const int nouter1 = T[nb].neighbor[(ne + 2) % 3], nouter2 = T[nb].neighbor[(ne + 1) % 3];
		T[c3].neighbor[1] = nouter1;
		T[c4].neighbor[1] = nouter2;
		if (T[nb].parent == nb) {
			T[nb].parent = -1;
		} else {
			T[nb] = T[c4];
			c4 = nb;
			m.trinum--;
		}
		relink(nouter1, nb, c3);
		relink(nouter2, nb, c4);
		T[c3].neighbor[0] = c4;
		T[c4].neighbor[2] = c3;
		// shade.cpp:416 compares the addresses of two different arrays: never equal
		T[c1].neighbor[2] = c4;
		T[c4].neighbor[0] = c1;
		T[c2].neighbor[0] = c3;
		T[c3].neighbor[2] = c2;
		out[2] = c3, out[3] = c4;
		return replaced;
	}
};

}  // namespace

void prepartition(Model &m) {
	// The refinement runs in a calloc'ed pool of the reference's size: never-used slots are all-zero like the
	// reference's BSS, but only the pages that get written are ever touched (filling a 236 MB vector took longer
	// than the refinement itself); the live slots move into m.tri at the end.
	const int limit = m.pool_limit();
	const size_t cap = (size_t)std::max(limit, m.trinum);
	radk_triangle *pool = (radk_triangle *)calloc(cap, sizeof(radk_triangle));
	if (!pool) return;	// out of memory: the mesh stays unrefined
#if defined(__linux__)
	{	// a multi-GB pool walked in priority order: ask for huge pages (advice only; the block stays calloc's)
		const uintptr_t a = ((uintptr_t)pool + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1), e = ((uintptr_t)pool + cap * sizeof(radk_triangle)) & ~(uintptr_t)((2u << 20) - 1);
		if (e > a) madvise((void *)a, e - a, MADV_HUGEPAGE);
	}
#endif
	memcpy(pool, m.tri.data(), (size_t)m.trinum * sizeof(radk_triangle));
	Splitter sp{m, pool};
	// The reference keeps its work list in a std::set<pair<area, slot>> and takes *rbegin(): largest area first, larger
	// slot on ties.  Almost every push made while a triangle is split carries HALF of an area that is already in the list
	// ((float)((double)area / 2.0)), i.e. it belongs behind everything of the area being processed: the list is a map
	// area -> slots whose largest bucket is loaded, sorted once (descending) and walked.  The rare push that does not
	// go behind (the co-split children of a neighbour that was never listed, e.g. a light) lands in a small ordered
	// set that is consulted at every pop (church: 82 such pushes, hall 100, sbox 412 - the byte-identical meshes of the ten
	// models cover it).  Same pop sequence as the set (pairs are unique: entries of one slot always
	// differ in area, fresh slots are pushed once), without a 23 M-entry heap whose sift-downs miss the cache.
	struct WorkList {
		std::map<float, std::vector<int>> later;	// areas below the loaded bucket
		std::vector<int> bucket;			// the loaded bucket, slots descending
		size_t bi = 0;
		float bucket_area = 0.0f;
		bool loaded = false;
		std::set<std::pair<float, int>> hot;		// pushes at or above the loaded bucket's area

		void push(float a, int slot) {
			if (loaded && !(a < bucket_area)) hot.insert(std::make_pair(a, slot));
			else later[a].push_back(slot);
		}
		bool pop(std::pair<float, int> &out) {
			if (bi >= bucket.size()) {
				loaded = false;
				if (!later.empty()) {
					auto last = std::prev(later.end());
					bucket_area = last->first;
					bucket.clear();
					bucket.swap(last->second);
					later.erase(last);
					std::sort(bucket.begin(), bucket.end(), std::greater<int>());
					bi = 0, loaded = true;
				}
			}
			const bool hb = loaded && bi < bucket.size(), hh = !hot.empty();
			if (!hb && !hh) return false;
			if (hh && (!hb || *hot.rbegin() > std::make_pair(bucket_area, bucket[bi]))) {
				out = *hot.rbegin();
				hot.erase(std::prev(hot.end()));
			} else {
				out = std::make_pair(bucket_area, bucket[bi++]);
			}
			return true;
		}
	} work;
	for (int i = 0; i < m.trinum; i++) {
		const radk_triangle &t = pool[i];
		if (t.area < 1.f || is_light(t)) continue;
		work.push(t.area, i);
	}
	const bool exhaustive = m.opts.prepartition == 2;	// synthetic scaling scenes only: keep going where the reference gives up
	std::pair<float, int> top;
	while (work.pop(top)) {		// the reference dereferences rbegin() of an empty set here (UB)
		const int slot = top.second;
		if (pool[slot].area < 1) { if (exhaustive) continue; break; }
		int child[4];
		if (sp.split(slot, child) < 0) { if (exhaustive && !sp.pool_full(4)) continue; break; }
		if (is_light(pool[slot])) {	// shade.cpp:580-590
			const radk_triangle &t = pool[slot];
			for (int i = 0; i < 4; i++) {
				if (child[i] < 0) continue;
				radk_triangle &ch = pool[child[i]];
				for (int k = 0; k < 3; k++) ch.Brgb[k] = t.Brgb[k];
				for (int v = 0; v < 3; v++) {
					for (int k = 0; k < 3; k++) ch.accB[v][k] = t.accB[v][k];
					ch.deltaB[v] = t.deltaB[v];
				}
			}
		}
		for (int i = 0; i < 4; i++)
			if (child[i] >= 0) work.push(pool[child[i]].area, child[i]);
	}
	m.tri.adopt(pool, (size_t)m.trinum, cap);	// the pool becomes the store (trimmed by the caller's shrink_to_fit)
}

}  // namespace radk
