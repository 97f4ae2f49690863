// Emitter and neighbour initialisation: restates InitRadiosity (reference src/rad.cpp:259-338).
//  * lights (any Brgb == 255): accB = Brgb * light_scale per vertex, deltaB = sum_v 100*accB/3
//  * neighbour links across exactly shared edges whose normals agree (cos > 0.85)
// The reference scans all O(n0^2) pairs; the writes it performs depend only on pairs that share a
// vertex bit-for-bit, so the scan is done here as a hash join on vertex coordinates and the
// surviving pairs are replayed in the reference's (i ascending, j ascending) order, which yields
// the same final links as the serial reference loop.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include "model.hpp"

namespace radk {
namespace {

inline float dot(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline bool same_point(const float *a, const float *b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }

struct VKey {
	uint32_t x, y, z;
	bool operator==(const VKey &o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VKeyHash {
	size_t operator()(const VKey &k) const {
		uint64_t h = k.x * 0x9E3779B97F4A7C15ull;
		h ^= (h >> 29) + k.y * 0xBF58476D1CE4E5B9ull;
		h ^= (h >> 31) + k.z * 0x94D049BB133111EBull;
		return (size_t)(h ^ (h >> 32));
	}
};
inline uint32_t key_bits(float f) {
	if (f == 0.0f) f = 0.0f;	// -0 == +0 under operator==
	uint32_t u;
	memcpy(&u, &f, 4);
	return u;
}

}  // namespace

void init_radiosity(Model &m) {
	float minarea = 100000000.0f, maxarea = 0.0f;
	const int n = m.trinum;
	std::unordered_map<VKey, std::vector<int>, VKeyHash> by_vertex;
	by_vertex.reserve((size_t)n * 2);
	for (int i = 0; i < n; i++)
		for (int v = 0; v < 3; v++) {
			const float *p = m.tri[i].p[v];
			if (p[0] != p[0] || p[1] != p[1] || p[2] != p[2]) continue;	// NaN never compares equal
			auto &bucket = by_vertex[VKey{key_bits(p[0]), key_bits(p[1]), key_bits(p[2])}];
			if (bucket.empty() || bucket.back() != i) bucket.push_back(i);
		}
	std::vector<int> cand;
	for (int i = 0; i < n; i++) {
		radk_triangle &t = m.tri[i];
		minarea = t.area < minarea ? t.area : minarea;
		maxarea = maxarea < t.area ? t.area : maxarea;
		if (is_light(t)) {
			t.deltaB[0] = t.deltaB[1] = t.deltaB[2] = 0.0f;
			const float scale = m.opts.light_scale;
			for (int v = 0; v < 3; v++)
				for (int c = 0; c < 3; c++) {
					t.accB[v][c] = t.Brgb[c] * scale;
					t.deltaB[c] += 100 * t.accB[v][c] / 3;
				}
		} else {
			memset(t.accB, 0, sizeof(t.accB));
			t.deltaB[0] = t.deltaB[1] = t.deltaB[2] = 0.0f;
		}
		// neighbour candidates: triangles sharing at least one vertex with i, ascending
		cand.clear();
		for (int v = 0; v < 3; v++) {
			const float *p = t.p[v];
			if (p[0] != p[0] || p[1] != p[1] || p[2] != p[2]) continue;
			auto it = by_vertex.find(VKey{key_bits(p[0]), key_bits(p[1]), key_bits(p[2])});
			if (it != by_vertex.end()) cand.insert(cand.end(), it->second.begin(), it->second.end());
		}
		std::sort(cand.begin(), cand.end());
		cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
		for (int j : cand) {
			if (j == i) continue;
			radk_triangle &u = m.tri[j];
			for (int v = 0; v < 3; v++)
				for (int w = 0; w < 3; w++) {
					if (!same_point(t.p[v], u.p[w])) continue;
					const float *tn = t.p[(v + 1) % 3];
					int slot;
					if (same_point(tn, u.p[(w + 1) % 3])) slot = w;			// same winding
					else if (same_point(tn, u.p[(w + 2) % 3])) slot = (w + 2) % 3;	// opposite winding
					else continue;
					float cosang = dot(t.n, u.n) / sqrtf(dot(t.n, t.n)) / sqrtf(dot(u.n, u.n));	// vector.h:40
					if (cosang > 0.85) {
						t.neighbor[v] = j;
						u.neighbor[slot] = i;
					}
				}
		}
	}
	m.min_area = minarea;
	m.max_area = maxarea;
	m.sample_area = m.opts.sample_area;
	if (m.sample_area < 0.0) m.sample_area = (float)((double)maxarea / 100.0);
}

}  // namespace radk
