// Octree over the input triangles: restates BuildTree / buildTree / CrossOver / Clip /
// InterpVector (reference src/raycast.cpp:24-233).  The tree is tiny (church: 8.7 k nodes) and
// built once, on the host; the device consumes a flattened copy (csrc/scene_upload.cu).
// Node order is DFS pre-order with children 0..7 (bit j of the child index = upper half on axis j),
// leaves when <= 10 triangles or depth 6; leaf lists are the parent's list filtered by the
// slab/clip overlap test, terminated by -1.
#include <cstring>
#include "model.hpp"

namespace radk {
namespace {

struct Builder {
	Model &m;
	int max_nodes;
	bool overflow = false;

	// raycast.cpp:24-60: does segment a-b touch the rectangle [g0,g1] in the (x,y) plane?
	static bool segment_touches_rect(const float *a, const float *b, const float *g0, const float *g1, int x, int y) {
		int ca = (a[x] < g0[x]) | (a[x] > g1[x]) << 1 | (a[y] < g0[y]) << 2 | (a[y] > g1[y]) << 3;
		int cb = (b[x] < g0[x]) | (b[x] > g1[x]) << 1 | (b[y] < g0[y]) << 2 | (b[y] > g1[y]) << 3;
		if (ca & cb) return false;
		if (!(ca | cb)) return true;
		float dy = b[y] - a[y], dx = b[x] - a[x];
		float side = dy * (g0[x] - a[x]) - dx * (g0[y] - a[y]);
		if (side == 0) return true;
		bool neg = side < 0;
		if (((dy * (g0[x] - a[x]) - dx * (g1[y] - a[y])) < 0) != neg) return true;
		if (((dy * (g1[x] - a[x]) - dx * (g0[y] - a[y])) < 0) != neg) return true;
		if (((dy * (g1[x] - a[x]) - dx * (g1[y] - a[y])) < 0) != neg) return true;
		return false;
	}

	// raycast.cpp:91-100: point on segment hi-lo whose coordinate d equals g
	static void point_at(const float *hi, const float *lo, float g, int d, float *out) {
		float span = hi[d] - lo[d];
		if (span == 0.0) {
			out[0] = hi[0], out[1] = hi[1], out[2] = hi[2];
		} else {
			float wa = (g - lo[d]) / span, wb = (hi[d] - g) / span;
			for (int k = 0; k < 3; k++) out[k] = wa * hi[k] + wb * lo[k];
		}
	}

	// raycast.cpp:109-155: approximate triangle / cube overlap by slicing with the cube's faces
	static bool overlaps(const radk_triangle &t, const float *g0, const float *g1) {
		const float *hi = t.p[0], *mid = t.p[1], *lo = t.p[2];
		float a[3], b[3];
		for (int x = 0; x < 3; x++) {
			// three-element sort by coordinate x (the order carries over between axes)
			if (hi[x] < mid[x]) { const float *s = hi; hi = mid; mid = s; }
			if (hi[x] < lo[x]) { const float *s = hi; hi = lo; lo = s; }
			if (mid[x] < lo[x]) { const float *s = mid; mid = lo; lo = s; }
			const int y = (x + 1) % 3, z = (x + 2) % 3;
			if (g1[x] < lo[x]) return false;
			point_at(hi, lo, g1[x], x, a);
			if (g1[x] < mid[x]) {
				point_at(mid, lo, g1[x], x, b);
				return segment_touches_rect(a, b, g0, g1, y, z);
			} else if (g1[x] <= hi[x]) {
				point_at(hi, mid, g1[x], x, b);
				return segment_touches_rect(a, b, g0, g1, y, z);
			}
			if (g0[x] > hi[x]) return false;
			point_at(hi, lo, g0[x], x, a);
			if (g0[x] > mid[x]) {
				point_at(hi, mid, g0[x], x, b);
				return segment_touches_rect(a, b, g0, g1, y, z);
			} else if (g0[x] >= lo[x]) {
				point_at(mid, lo, g0[x], x, b);
				return segment_touches_rect(a, b, g0, g1, y, z);
			}
		}
		return true;
	}

	// raycast.cpp:168-213
	int build(const std::vector<int> &ids, int level, const float *g0, const float *g1) {
		const int index = (int)m.nodes.size();
		m.nodes.emplace_back();
		if ((int)m.nodes.size() >= max_nodes) { overflow = true; return index; }
		{
			radk_tree_node &nd = m.nodes[index];
			for (int k = 0; k < 3; k++) nd.g0[k] = g0[k], nd.g1[k] = g1[k];
			nd.list = -1;
			for (int k = 0; k < 8; k++) nd.sub[k] = -1;
		}
		if ((int)ids.size() <= kMaxTriPerNode || level >= kMaxLevel) {
			if (g1[0] - g0[0] < m.min_grid_len) m.min_grid_len = g1[0] - g0[0];
			m.nodes[index].list = (int)m.list.size();
			m.list.insert(m.list.end(), ids.begin(), ids.end());
			m.list.push_back(-1);
			return index;
		}
		float mid[3];
		for (int k = 0; k < 3; k++) mid[k] = (float)(0.5 * (double)g0[k] + 0.5 * (double)g1[k]);	// vector.h:23-26 with double 0.5
		for (int i = 0; i < 8 && !overflow; i++) {
			float q0[3] = {g0[0], g0[1], g0[2]}, q1[3] = {g1[0], g1[1], g1[2]};
			for (int j = 0; j < 3; j++) {
				if ((i >> j) & 1) q0[j] = mid[j];
				else q1[j] = mid[j];
			}
			std::vector<int> inside;
			inside.reserve(ids.size());
			for (int id : ids)
				if (overlaps(m.tri[id], q0, q1)) inside.push_back(id);
			int child = build(inside, level + 1, q0, q1);
			m.nodes[index].sub[i] = child;
		}
		return index;
	}
};

}  // namespace

int build_octree(Model &m) {
	m.nodes.clear();
	m.list.clear();
	m.min_grid_len = m.G1[0] - m.G0[0];
	std::vector<int> all(m.trinum);
	for (int i = 0; i < m.trinum; i++) all[i] = i;
	Builder b{m, m.opts.max_tree_nodes > 0 ? m.opts.max_tree_nodes : kRefMaxTreeNode};
	b.build(all, 0, m.G0, m.G1);
	if (b.overflow) {
		set_error("#TreeNode exceeded limitation (reference: exit(1), raycast.cpp:175-178)");
		return RADK_ERR_LIMIT;
	}
	m.grid_num = (int)((m.G1[0] - m.G0[0]) / m.min_grid_len);
	return RADK_OK;
}

}  // namespace radk
