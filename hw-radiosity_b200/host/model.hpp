// Host front end of the B200 radiosity path: scene model held in host memory.
// Mirrors the reference's global state (TriStore/trinum/G0/G1, rad.cpp:33-36, and the octree
// file-statics raycast.cpp:9-16) as one object instead of globals.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "radk.h"

namespace radk {

static_assert(sizeof(radk_triangle) == 236, "radk_triangle must match the reference Triangle");
static_assert(sizeof(radk_tree_node) == 60, "radk_tree_node must match the reference TreeNode");

constexpr double kPi = 3.1415926535897932;	// rad.h:5
constexpr int kRefMaxTri = 1000000;		// rad.h:9
constexpr int kRefMaxTreeNode = 20000;		// rad.h:11
constexpr int kMaxTriPerNode = 10;		// rad.h:12
constexpr int kMaxLevel = 6;			// rad.h:7

// The slot store (the reference's static TriStore[], rad.cpp:33): malloc-backed so that PrePartition can refine in a
// calloc'ed pool (never-written slots stay untouched zero pages, like the reference's BSS) and simply hand the pool
// over - copying 5 GB of records into a vector was a fifth of the set-up of an 11.6 M-receiver scene.
class TriBuf {
public:
	TriBuf() = default;
	TriBuf(const TriBuf &) = delete;
	TriBuf &operator=(const TriBuf &) = delete;
	~TriBuf() { free(p_); }
	radk_triangle *data() { return p_; }
	const radk_triangle *data() const { return p_; }
	size_t size() const { return n_; }
	radk_triangle &operator[](size_t i) { return p_[i]; }
	const radk_triangle &operator[](size_t i) const { return p_[i]; }
	void clear() { n_ = 0; }
	bool reserve(size_t cap) {
		if (cap <= cap_) return true;
		void *q = realloc(p_, cap * sizeof(radk_triangle));
		if (!q) return false;
		p_ = (radk_triangle *)q, cap_ = cap;
		return true;
	}
	void push_back(const radk_triangle &t) {
		if (n_ == cap_ && !reserve(cap_ ? 2 * cap_ : 1024)) abort();
		p_[n_++] = t;
	}
	void assign(const radk_triangle *b, const radk_triangle *e) {
		const size_t n = (size_t)(e - b);
		if (!reserve(n)) abort();
		memcpy(p_, b, n * sizeof(radk_triangle));
		n_ = n;
	}
	void resize(size_t n) {	// new slots are zero, like the reference's BSS
		if (n > cap_ && !reserve(n)) abort();
		if (n > n_) memset((void *)(p_ + n_), 0, (n - n_) * sizeof(radk_triangle));
		n_ = n;
	}
	void shrink_to_fit() {
		if (n_ < cap_ && n_ > 0) {
			void *q = realloc(p_, n_ * sizeof(radk_triangle));
			if (q) p_ = (radk_triangle *)q, cap_ = n_;
		}
	}
	void adopt(radk_triangle *p, size_t n, size_t cap) { free(p_); p_ = p, n_ = n, cap_ = cap; }	// takes a malloc / calloc block
private:
	radk_triangle *p_ = nullptr;
	size_t n_ = 0, cap_ = 0;
};

struct Model {
	TriBuf tri;				// live slots [0, trinum)
	int trinum = 0;
	int n0 = 0;
	float G0[3] = {0, 0, 0}, G1[3] = {0, 0, 0};
	std::vector<radk_tree_node> nodes;
	std::vector<int> list;
	float min_grid_len = 0;
	int grid_num = 0;
	float min_area = 0, max_area = 0, sample_area = -1;
	radk_model_opts opts{};
	// snapshot writer state (fin_writer.cpp): the vertex/normal text of a frozen mesh never changes, so it is
	// formatted once; a snapshot then only needs one colour line per leaf
	std::string geom_text;			// concatenated "x y z nx ny nz\n" x3 chunks of all leaves, slot order
	std::vector<size_t> geom_off;		// chunk boundaries (leaves + 1 entries)
	std::vector<int> geom_slot;		// slot of each leaf
	bool geom_valid = false;
	std::thread writer;			// at most one snapshot file being written in the background
	int writer_rc = 0;

	int pool_limit() const {
		int mt = opts.max_tri > 0 ? opts.max_tri : kRefMaxTri;
		return mt < opts.tri_limit ? mt : opts.tri_limit;
	}
};

inline bool is_light(const radk_triangle &t) {	// rad.h:47-49
	return t.Brgb[0] == 255 || t.Brgb[1] == 255 || t.Brgb[2] == 255;
}

void set_error(const std::string &msg);

// tri_reader.cpp  (rad.cpp:47-111)
int read_triangles(const char *text, size_t len, Model &m);
// octree.cpp      (raycast.cpp:24-233)
int build_octree(Model &m);
// emitters.cpp    (rad.cpp:259-338)
void init_radiosity(Model &m);
// prepartition.cpp (shade.cpp:171-211, 311-431, 556-598)
void prepartition(Model &m);
// fin_writer.cpp  (rad.cpp:117-221 + float2string.h)
void format_fin(const Model &m, std::string &out);
void build_geometry_cache(Model &m);
// text of a snapshot from per-slot colours rgb[trinum][3] (logical slots ignored) and the cached geometry
void format_fin_colors(const Model &m, const int *rgb, std::string &out);
int float_to_text(float v, char *buf);	// fp::tostring<float> equivalent; returns length, 0 = nothing written

}  // namespace radk
