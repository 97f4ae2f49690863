// .tri reader: restates ReadTriangle (reference src/rad.cpp:47-111) for an in-memory text.
// Per triangle: "Triangle" token, 3 int Frgb, 3 int Brgb, 3x3 float vertices; derived centre,
// normalised normal, area and the three inward side planes; scene cube G0/G1 from the bounds.
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "model.hpp"

namespace radk {
namespace {

struct Cursor {
	const char *s, *end;
	void skip_ws() { while (s < end && (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r' || *s == '\v' || *s == '\f')) ++s; }
	bool token() {			// fscanf("%s")
		skip_ws();
		if (s >= end) return false;
		while (s < end && !(*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r' || *s == '\v' || *s == '\f')) ++s;
		return true;
	}
	bool integer(int &v) {		// fscanf("%d")
		skip_ws();
		if (s >= end) return false;
		char *e;
		long x = strtol(s, &e, 10);
		if (e == s) return false;
		v = (int)x; s = e;
		return true;
	}
	bool real(float &v) {		// fscanf("%f")
		skip_ws();
		if (s >= end) return false;
		char *e;
		v = strtof(s, &e);
		if (e == s) return false;
		s = e;
		return true;
	}
};

inline float dot(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void cross(const float *a, const float *b, float *o) {	// vector.h:71-75
	o[0] = a[1] * b[2] - b[1] * a[2];
	o[1] = a[2] * b[0] - b[2] * a[0];
	o[2] = a[0] * b[1] - b[0] * a[1];
}

}  // namespace

// Geometry that ReadTriangle derives for one input triangle (rad.cpp:76-88).
static void derive_geometry(radk_triangle &t) {
	for (int k = 0; k < 3; k++)	// CalCenter, vector.h:54-57: float sum, double /3.0
		t.c[k] = (float)((double)(t.p[0][k] + t.p[1][k] + t.p[2][k]) / 3.0);
	float e0[3], e1[3];
	for (int k = 0; k < 3; k++) e0[k] = t.p[1][k] - t.p[0][k], e1[k] = t.p[2][k] - t.p[0][k];
	cross(e0, e1, t.n);		// CalArea, vector.h:58-62
	t.area = sqrtf(dot(t.n, t.n)) / 2;
	float len = sqrtf(dot(t.n, t.n));	// normalize, vector.h:76-79
	t.n[0] /= len, t.n[1] /= len, t.n[2] /= len;
	for (int i = 0; i < 3; i++) {	// inward side planes
		float edge[3], sn[3];
		const float *a = t.p[i], *b = t.p[(i + 1) % 3];
		for (int k = 0; k < 3; k++) edge[k] = b[k] - a[k];
		cross(t.n, edge, sn);
		t.se[i][0] = sn[0], t.se[i][1] = sn[1], t.se[i][2] = sn[2];
		t.se[i][3] = -dot(a, sn);
	}
}

// .poly front end (the reference itself cannot read its model/*.poly files):
// records "Fr Fg Fb Br Bg Bb N" followed by N vertex lines; an N-gon is fanned from its first vertex
// (a b c d -> a b c, a c d), which is how the reference's sibling .tri files were made from the .poly ones.
static bool poly_to_tri_text(const std::string &in, std::string &out) {
	Cursor cur{in.c_str(), in.c_str() + in.size()};
	char line[256];
	for (;;) {
		cur.skip_ws();
		if (cur.s >= cur.end) break;
		int col[6], n;
		for (int k = 0; k < 6; k++)
			if (!cur.integer(col[k])) return false;
		if (!cur.integer(n) || n < 3 || n > 4096) return false;
		std::vector<std::string> v((size_t)n);
		for (int i = 0; i < n; i++) {		// vertex coordinates are passed through as text: no re-rounding
			for (int k = 0; k < 3; k++) {
				cur.skip_ws();
				const char *b = cur.s;
				float dummy;
				if (!cur.real(dummy)) return false;
				if (k) v[i] += ' ';
				v[i].append(b, cur.s);
			}
		}
		for (int i = 1; i + 1 < n; i++) {
			snprintf(line, sizeof(line), "Triangle\n%d %d %d %d %d %d\n", col[0], col[1], col[2], col[3], col[4], col[5]);
			out += line;
			out += v[0] + "\n" + v[i] + "\n" + v[i + 1] + "\n";
		}
	}
	return true;
}

int read_triangles(const char *text, size_t len, Model &m) {
	// strtof/strtol need a terminated buffer
	std::string owned(text, len);
	{	// a .poly file starts with a number, a .tri file with the "Triangle" keyword
		size_t i = 0;
		while (i < owned.size() && (owned[i] == ' ' || owned[i] == '\t' || owned[i] == '\n' || owned[i] == '\r')) i++;
		if (i < owned.size() && (owned[i] == '-' || owned[i] == '+' || (owned[i] >= '0' && owned[i] <= '9'))) {
			std::string tri;
			if (!poly_to_tri_text(owned, tri)) { set_error("malformed .poly input"); return RADK_ERR_IO; }
			owned.swap(tri);
		}
	}
	Cursor cur{owned.c_str(), owned.c_str() + owned.size()};
	float lo[3] = {1000000.0f, 1000000.0f, 1000000.0f}, hi[3] = {-1000000.0f, -1000000.0f, -1000000.0f};
	const int cap = m.opts.max_tri > 0 ? m.opts.max_tri : kRefMaxTri;
	m.tri.clear();
	m.trinum = 0;
	for (;;) {
		if (!cur.token()) break;
		cur.skip_ws();
		if (cur.s >= cur.end) break;	// reference: "|| feof(fp)" after the keyword
		if (m.trinum >= cap) { set_error("model has more triangles than max_tri"); return RADK_ERR_LIMIT; }
		radk_triangle t;
		memset(&t, 0, sizeof(t));
		bool ok = true;
		for (int k = 0; k < 3; k++) ok = ok && cur.integer(t.Frgb[k]);
		for (int k = 0; k < 3; k++) ok = ok && cur.integer(t.Brgb[k]);
		for (int v = 0; v < 3; v++)
			for (int k = 0; k < 3; k++) ok = ok && cur.real(t.p[v][k]);
		if (!ok) { set_error("malformed .tri input (the reference asserts here, vector.h:43-44)"); return RADK_ERR_IO; }
		derive_geometry(t);
		for (int v = 0; v < 3; v++)
			for (int k = 0; k < 3; k++) {
				lo[k] = t.p[v][k] < lo[k] ? t.p[v][k] : lo[k];
				hi[k] = hi[k] < t.p[v][k] ? t.p[v][k] : hi[k];
			}
		t.parent = m.trinum;
		t.neighbor[0] = t.neighbor[1] = t.neighbor[2] = -1;
		m.tri.push_back(t);
		m.trinum++;
	}
	m.n0 = m.trinum;
	float ext = 0;
	for (int k = 0; k < 3; k++) ext = ext < (hi[k] - lo[k]) ? (hi[k] - lo[k]) : ext;
	for (int k = 0; k < 3; k++) {	// rad.cpp:109-110: float - double PI, float + float + double 2PI
		m.G0[k] = (float)((double)lo[k] - kPi);
		m.G1[k] = (float)((double)(m.G0[k] + ext) + 2 * kPi);
	}
	return RADK_OK;
}

}  // namespace radk
