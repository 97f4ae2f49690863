// Host-side flattening of a scene into the arrays the kernels read (pure C++, no CUDA calls):
//  * the octree becomes a dense cell->leaf table + per-node float4 records + the leaf lists,
//  * occluders TriStore[0..n0) become 5 float4 each,
//  * receivers (slots with parent >= 0) are sorted by (original parent, Morton code), dealt to the ranks
//    in blocks of 512, and split into float4 planes; vertices shared by receivers of one original are merged.
#pragma once
#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>
#include <unordered_map>
#include <thread>
#include <vector>

#include "radk.h"
#include "rt_core.cuh"

namespace radk {

struct FlatScene {
	// octree
	std::vector<int> cell_leaf;
	std::vector<float4> leaf_lo, leaf_hi;	// host-side views; the device gets them interleaved as leaf_box
	std::vector<float4> leaf_box;
	std::vector<int> list;
	std::vector<int4> list4;
	std::vector<int> leaf_list4, tri_leaf_begin, tri_leaves;
	std::vector<float4> occ;
	// receivers of this shard
	int R = 0, U = 0, leaves_total = 0;
	std::vector<int> recv_slot, recv_parent;
	std::vector<float4> geo[3], st[3];
	std::vector<int> slot_to_recv;
	std::vector<float4> par_n, par_fk;
	std::vector<int4> vid;
	std::vector<float4> uv;
	std::vector<int> uv_leaf;	// leaf that contains the vertex = first leaf of every ray that starts there
	std::string error;

	DevTree host_tree(const radk_scene *s) const {	// the same view the kernels get, over host memory
		DevTree T;
		T.cell_leaf = cell_leaf.data(), T.leaf_box = leaf_box.data();
		T.list = list.data(), T.occ = occ.data(), T.nocc = s->n0;
		T.list4 = list4.data(), T.leaf_list4 = leaf_list4.data();
		T.tri_leaf_begin = tri_leaf_begin.data(), T.tri_leaves = tri_leaves.data();
		T.nnodes = (int)leaf_lo.size();
		for (int k = 0; k < 3; k++) T.G0[k] = s->G0[k], T.G1[k] = s->G1[k];
		T.min_grid_len = s->min_grid_len, T.grid_num = s->grid_num;
		T.inv_min_grid_len = 1.0 / (double)s->min_grid_len;
		T.inv_min_grid_len_f = (float)T.inv_min_grid_len;
		T.cell_eps = 4e-7f * (float)(s->grid_num + 1) + 1e-4f;
		return T;
	}
};

namespace flat_detail {
inline uint32_t spread3(uint32_t x) {	// 10 bits -> every third bit
	x &= 0x3ff;
	x = (x | (x << 16)) & 0x030000ff;
	x = (x | (x << 8)) & 0x0300f00f;
	x = (x | (x << 4)) & 0x030c30c3;
	x = (x | (x << 2)) & 0x09249249;
	return x;
}
struct VKey {
	uint32_t x, y, z;
	int parent;
	bool operator==(const VKey &o) const { return x == o.x && y == o.y && z == o.z && parent == o.parent; }
};
struct VKeyHash {
	size_t operator()(const VKey &k) const {
		uint64_t h = k.x * 0x9E3779B97F4A7C15ull;
		h ^= (h >> 29) + k.y * 0xBF58476D1CE4E5B9ull;
		h ^= (h >> 31) + k.z * 0x94D049BB133111EBull;
		h ^= (h >> 27) + (uint32_t)k.parent * 0xD6E8FEB86659FD93ull;
		return (size_t)(h ^ (h >> 32));
	}
};
inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float ibits(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float4 f4(float a, float b, float c, float d) { float4 r; r.x = a, r.y = b, r.z = c, r.w = d; return r; }
inline int4 i4(int a, int b, int c, int d) { int4 r; r.x = a, r.y = b, r.z = c, r.w = d; return r; }
}  // namespace flat_detail

// Set-up passes over millions of 236-byte records are memory bound and independent per record: a few host threads share
// them (the results do not depend on the number of threads).
template <typename Fn> inline void flat_parallel(size_t n, Fn fn) {
	const unsigned hw = std::thread::hardware_concurrency();
	const size_t nt = n < 200000 ? 1 : std::min<size_t>(8, std::max(1u, hw));
	if (nt <= 1) { fn((size_t)0, n, 0); return; }
	std::vector<std::thread> th;
	const size_t chunk = (n + nt - 1) / nt;
	for (size_t t = 0; t < nt; t++) th.emplace_back([=] { fn(std::min(n, t * chunk), std::min(n, (t + 1) * chunk), (int)t); });
	for (auto &x : th) x.join();
}

// Returns a RADK_* code; on failure F.error holds the message.
inline int flatten_scene(const radk_scene *s, int rank, int world, FlatScene &F) {
	using namespace flat_detail;
	// ---- octree: dense cell table by running the reference descent (raycast.cpp:266-278) per cell
	const int g = s->grid_num, dim = g + 1;
	if (g <= 0 || (long long)dim * dim * dim > (1ll << 28)) { F.error = "grid_num out of range"; return RADK_ERR_ARG; }
	if (s->nnodes <= 0) { F.error = "empty octree"; return RADK_ERR_ARG; }
	std::vector<int> depth(s->nnodes, 0);
	F.cell_leaf.assign((size_t)dim * dim * dim, 0);
	for (int z = 0; z < dim; z++)
		for (int y = 0; y < dim; y++)
			for (int x = 0; x < dim; x++) {
				int gn[3] = {x, y, z}, node = 0, gsize = g, d = 0;
				while (s->nodes[node].list < 0) {
					gsize >>= 1;
					int idx = 0;
					for (int i = 0; i < 3; i++)
						if (gn[i] >= gsize) idx |= 1 << i, gn[i] -= gsize;
					node = s->nodes[node].sub[idx];
					d++;
					if (node < 0 || node >= s->nnodes || d > 64) { F.error = "malformed octree"; return RADK_ERR_ARG; }
				}
				depth[node] = d;
				F.cell_leaf[((size_t)z * dim + y) * dim + x] = node;
			}
	F.leaf_lo.resize(s->nnodes), F.leaf_hi.resize(s->nnodes);
	for (int i = 0; i < s->nnodes; i++) {
		const radk_tree_node &nd = s->nodes[i];
		F.leaf_lo[i] = f4(nd.g0[0], nd.g0[1], nd.g0[2], ibits(nd.list));
		F.leaf_hi[i] = f4(nd.g1[0], nd.g1[1], nd.g1[2], ibits(depth[i]));
	}
	F.leaf_box.resize(2 * (size_t)s->nnodes);
	for (int i = 0; i < s->nnodes; i++) F.leaf_box[2 * (size_t)i] = F.leaf_lo[i], F.leaf_box[2 * (size_t)i + 1] = F.leaf_hi[i];
	F.list.assign(s->list, s->list + s->nlist);
	// padded copies of the leaf lists + inverted index triangle -> leaves
	F.leaf_list4.assign(s->nnodes, -1);
	F.list4.clear();
	std::vector<std::vector<int>> leaves_of(s->n0);
	for (int i = 0; i < s->nnodes; i++) {
		const int off = s->nodes[i].list;
		if (off < 0) continue;
		if (off >= s->nlist) { F.error = "leaf list offset out of range"; return RADK_ERR_ARG; }
		F.leaf_list4[i] = (int)F.list4.size();
		// any-hit order: occluders whose side planes are all zero first (cheapest test, most likely to hit)
		std::vector<int> planes, others;
		for (int k = off;; k++) {
			if (k >= s->nlist) { F.error = "unterminated leaf list"; return RADK_ERR_ARG; }
			const int idx = s->list[k];
			if (idx < 0) break;
			if (idx >= s->n0 || idx >= kPlaneFlag) { F.error = "leaf list references a non-input triangle"; return RADK_ERR_ARG; }
			if (leaves_of[idx].empty() || leaves_of[idx].back() != i) leaves_of[idx].push_back(i);
			bool zero_se = true;
			for (int a = 0; a < 3; a++)
				for (int b = 0; b < 4; b++) zero_se = zero_se && s->tri[idx].se[a][b] == 0.0f;
			(zero_se ? planes : others).push_back(zero_se ? (idx | kPlaneFlag) : idx);
		}
		planes.insert(planes.end(), others.begin(), others.end());
		planes.push_back(-1);
		while (planes.size() % 4) planes.push_back(-1);
		for (size_t k = 0; k < planes.size(); k += 4) F.list4.push_back(i4(planes[k], planes[k + 1], planes[k + 2], planes[k + 3]));
	}
	F.tri_leaf_begin.assign(s->n0 + 1, 0);
	F.tri_leaves.clear();
	for (int i = 0; i < s->n0; i++) {
		F.tri_leaf_begin[i] = (int)F.tri_leaves.size();
		F.tri_leaves.insert(F.tri_leaves.end(), leaves_of[i].begin(), leaves_of[i].end());
	}
	F.tri_leaf_begin[s->n0] = (int)F.tri_leaves.size();
	F.occ.resize((size_t)s->n0 * (OCC_PLANE + OCC_SIDE));
	for (int i = 0; i < s->n0; i++) {
		const radk_triangle &t = s->tri[i];
		float4 *q = &F.occ[(size_t)i * OCC_PLANE], *sd = &F.occ[(size_t)s->n0 * OCC_PLANE + (size_t)i * OCC_SIDE];
		q[0] = f4(t.n[0], t.n[1], t.n[2], t.c[0]);
		q[1] = f4(t.c[1], t.c[2], t.se[0][0], t.se[0][1]);
		sd[0] = f4(t.se[0][2], t.se[0][3], t.se[1][0], t.se[1][1]);
		sd[1] = f4(t.se[1][2], t.se[1][3], t.se[2][0], t.se[2][1]);
		sd[2] = f4(t.se[2][2], t.se[2][3], 0.f, 0.f);
	}

	// ---- receivers: leaves sorted by (parent, Morton code of the centroid), then sharded
	std::vector<int> leaves;
	leaves.reserve(s->trinum / 2 + 1);
	for (int i = 0; i < s->trinum; i++)
		if (s->tri[i].parent >= 0) {
			if (s->tri[i].parent >= s->n0) { F.error = "parent index outside the input triangles"; return RADK_ERR_ARG; }
			leaves.push_back(i);
		}
	const float ext = s->G1[0] - s->G0[0];
	auto morton = [&](const float *p) {
		uint32_t q[3];
		for (int k = 0; k < 3; k++) {
			float f = ext > 0 ? (p[k] - s->G0[k]) / ext : 0.f;
			f = f < 0.f ? 0.f : (f > 1.f ? 1.f : f);
			q[k] = (uint32_t)(f * 1023.0f);
		}
		return spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
	};
	std::vector<uint64_t> lkey(leaves.size());
	flat_parallel(leaves.size(), [&](size_t b, size_t e, int) {
		for (size_t i = b; i < e; i++) {
			const radk_triangle &t = s->tri[leaves[i]];
			const float cen[3] = {(t.p[0][0] + t.p[1][0] + t.p[2][0]) / 3, (t.p[0][1] + t.p[1][1] + t.p[2][1]) / 3,
					      (t.p[0][2] + t.p[1][2] + t.p[2][2]) / 3};
			lkey[i] = ((uint64_t)(uint32_t)t.parent << 32) | morton(cen);
		}
	});
	// (key, position in `leaves`) pairs sorted directly (leaves is ascending, so the position breaks ties like the slot
	// does): an index sort through a comparator that chases two arrays took three times as long on 11.6 M receivers
	std::vector<std::pair<uint64_t, int>> lsort(leaves.size());
	for (size_t i = 0; i < leaves.size(); i++) lsort[i] = std::make_pair(lkey[i], (int)i);
	std::sort(lsort.begin(), lsort.end());
	std::vector<int> order(leaves.size());
	for (size_t i = 0; i < leaves.size(); i++) order[i] = lsort[i].second;
	std::vector<std::pair<uint64_t, int>>().swap(lsort);
	F.leaves_total = (int)leaves.size();
	if (world < 1) world = 1;
	// block-cyclic sharding of the sorted receivers: blocks of kShardBlock neighbours are dealt round-robin, so
	// every rank gets an even sample of the scene (contiguous halves differ a lot in how far their rays travel)
	constexpr int kShardBlock = 512;
	std::vector<int> mine;
	mine.reserve(leaves.size() / world + kShardBlock);
	for (size_t i = 0; i < leaves.size(); i++)
		if ((int)((i / kShardBlock) % (size_t)world) == rank) mine.push_back((int)i);
	const int R = (int)mine.size();
	F.R = R;
	F.recv_slot.resize(R);
	F.recv_parent.resize(R);
	for (int v = 0; v < 3; v++) F.geo[v].resize(R), F.st[v].resize(R);
	F.slot_to_recv.assign(s->trinum, -1);
	F.par_n.assign(s->n0, f4(0, 0, 0, 0));
	F.par_fk.assign(s->n0, f4(0, 0, 0, 0));
	// per-original tables cover every original (also those whose receivers live on other ranks):
	// a shooter owned elsewhere still needs its normal here.
	const float kref = (float)(0.5 / 255.0);	// shade.cpp:246
	{
		// every thread notes, per original, the first normal / reflectance it meets in its range and checks the rest of
		// its range against it; the per-thread tables are then checked against each other
		struct Seen { std::vector<char> seen; std::vector<float4> n, fk; bool bad = false; };
		std::vector<Seen> part(8);
		for (auto &q : part) q.seen.assign(s->n0, 0), q.n.assign(s->n0, f4(0, 0, 0, 0)), q.fk.assign(s->n0, f4(0, 0, 0, 0));
		flat_parallel(leaves.size(), [&](size_t b, size_t e, int tid) {
			Seen &q = part[tid];
			for (size_t i = b; i < e; i++) {
				const radk_triangle &t = s->tri[leaves[i]];
				const float nlen = sqrtf(t.n[0] * t.n[0] + t.n[1] * t.n[1] + t.n[2] * t.n[2]);
				const float4 pn = f4(t.n[0], t.n[1], t.n[2], nlen);
				const float4 fk = f4(t.Frgb[0] * kref, t.Frgb[1] * kref, t.Frgb[2] * kref, 0.f);
				if (!q.seen[t.parent]) q.seen[t.parent] = 1, q.n[t.parent] = pn, q.fk[t.parent] = fk;
				else if (memcmp(&q.n[t.parent], &pn, 12) != 0 || memcmp(&q.fk[t.parent], &fk, 12) != 0) q.bad = true;
			}
		});
		std::vector<char> par_seen(s->n0, 0);
		bool bad = false;
		for (auto &q : part) {
			bad = bad || q.bad;
			for (int o = 0; o < s->n0; o++) {
				if (!q.seen[o]) continue;
				if (!par_seen[o]) par_seen[o] = 1, F.par_n[o] = q.n[o], F.par_fk[o] = q.fk[o];
				else if (memcmp(&F.par_n[o], &q.n[o], 12) != 0 || memcmp(&F.par_fk[o], &q.fk[o], 12) != 0) bad = true;
			}
		}
		if (bad) {
			F.error = "receivers of one original disagree on normal/reflectance: not a mesh produced by PrePartitionTriangles";
			return RADK_ERR_STATE;
		}
	}
	flat_parallel((size_t)R, [&](size_t b, size_t e, int) {
		for (size_t r = b; r < e; r++) {
			const int slot = leaves[order[mine[r]]];
			const radk_triangle &t = s->tri[slot];
			F.recv_slot[r] = slot;
			F.recv_parent[r] = t.parent;
			F.slot_to_recv[slot] = (int)r;
			F.geo[0][r] = f4(t.p[0][0], t.p[0][1], t.p[0][2], ibits(t.parent));
			F.geo[1][r] = f4(t.p[1][0], t.p[1][1], t.p[1][2], ibits(slot));
			F.geo[2][r] = f4(t.p[2][0], t.p[2][1], t.p[2][2], t.area);
			for (int v = 0; v < 3; v++) F.st[v][r] = f4(t.accB[v][0], t.accB[v][1], t.accB[v][2], t.deltaB[v]);
		}
	});

	// ---- unique (vertex, parent) pairs of this shard, sorted by (parent, Morton)
	// open-addressing table of vertex ids (keys compared through ukeys[]): ids are handed out in order of first
	// appearance, like the node-based map this replaces (which spent its time in malloc and pointer chasing)
	size_t tcap = 1024;
	while (tcap < (size_t)R * 4) tcap <<= 1;	// load factor <= 0.75 even if no vertex is shared
	std::vector<int> table(tcap, -1);
	const size_t tmask = tcap - 1;
	std::vector<VKey> ukeys;
	std::vector<float4> upos;
	ukeys.reserve((size_t)R), upos.reserve((size_t)R);
	F.vid.resize(R);
	const VKeyHash vhash;
	for (int r = 0; r < R; r++) {
		int ids[3];
		for (int v = 0; v < 3; v++) {
			const float4 &gp = F.geo[v][r];
			// exact bit patterns: p enters the arithmetic by value, so -0/+0 are not merged
			const VKey k{fbits(gp.x), fbits(gp.y), fbits(gp.z), F.recv_parent[r]};
			size_t h = vhash(k) & tmask;
			while (table[h] >= 0 && !(ukeys[table[h]] == k)) h = (h + 1) & tmask;
			if (table[h] < 0) {
				table[h] = (int)ukeys.size();
				ukeys.push_back(k);
				upos.push_back(f4(gp.x, gp.y, gp.z, ibits(F.recv_parent[r])));
			}
			ids[v] = table[h];
		}
		F.vid[r] = i4(ids[0], ids[1], ids[2], F.recv_parent[r]);
	}
	std::vector<int>().swap(table);
	const int U = (int)ukeys.size();
	F.U = U;
	std::vector<uint64_t> ukey2(U);
	// order: start leaf of the vertex's rays (every ray begins in the leaf that contains the vertex, so a warp's
	// lanes scan the same triangle list in their first - often only - leaf visit), then Morton code inside the leaf
	{
		const DevTree T = F.host_tree(s);
		for (int u = 0; u < U; u++) {
			const float p[3] = {upos[u].x, upos[u].y, upos[u].z};
			const int leaf = leaf_of<false>(T, p[0], p[1], p[2], nullptr);
			ukey2[u] = ((uint64_t)(uint32_t)(leaf + 1) << 32) | morton(p);
		}
	}
	std::vector<int> uorder(U), unew(U);
	{
		std::vector<std::pair<uint64_t, int>> usort(U);
		for (int u = 0; u < U; u++) usort[u] = std::make_pair(ukey2[u], u);
		std::sort(usort.begin(), usort.end());
		for (int u = 0; u < U; u++) uorder[u] = usort[u].second;
	}
	F.uv.resize(U);
	for (int i = 0; i < U; i++) unew[uorder[i]] = i, F.uv[i] = upos[uorder[i]];
	{
		const DevTree T = F.host_tree(s);
		F.uv_leaf.resize(U);
		for (int i = 0; i < U; i++) F.uv_leaf[i] = leaf_of<false>(T, F.uv[i].x, F.uv[i].y, F.uv[i].z, nullptr);
	}
	for (int r = 0; r < R; r++) F.vid[r] = i4(unew[F.vid[r].x], unew[F.vid[r].y], unew[F.vid[r].z], F.vid[r].w);
	return RADK_OK;
}

}  // namespace radk
