// Result writer: produces the text of PrintOut (reference src/rad.cpp:117-221, new format):
//   "Triangle \n"  "R G B 0 0 0\n"  then per vertex "x y z nx ny nz\n"
// for every non-logical slot in slot order.  Colour = trunc(clamp(mean of the three vertex accB,
// 0, 255)), NaN -> 0.  Numbers are printed with the same digit generation as the reference's
// vendored float2string.h (truncating, at most 8 fraction digits, trailing zeros trimmed;
// NaN/Inf/subnormal/huge values write nothing, so the previously printed token is repeated).
#include <cmath>
#include <cstdio>
#include <cstring>
#include "model.hpp"

namespace radk {

int float_to_text(float v, char *out) {
	uint32_t bits;
	memcpy(&bits, &v, 4);
	const bool neg = bits >> 31;
	const int biased = (bits >> 23) & 0xff;
	const uint32_t frac_bits = bits & 0x7fffffu;
	char *d = out;
	if (biased == 0xff) return 0;			// inf / nan: nothing written
	if (biased == 0) {
		if (frac_bits) return 0;		// subnormal: nothing written
		if (neg) *d++ = '-';
		*d++ = '0', *d++ = '.', *d++ = '0', *d = 0;
		return (int)(d - out);
	}
	const int sh = biased - 127 - 23;		// value = mant * 2^sh
	const uint64_t mant = frac_bits | 0x800000u;
	if (sh >= 64 - 23) return 0;			// integer part would not fit 64 bits: nothing written
	if (neg) *d++ = '-';
	// integer part
	uint64_t ip = sh > 0 ? mant << sh : (sh < -23 ? 0 : mant >> (-sh));
	if (ip == 0) {
		*d++ = '0';
	} else {
		char tmp[24];
		int n = 0;
		while (ip) tmp[n++] = (char)('0' + ip % 10), ip /= 10;
		while (n) *d++ = tmp[--n];
	}
	*d++ = '.';
	// fraction: 4.60 fixed point, eight truncated decimal digits, trailing zeros dropped
	const int up = 60 + sh;
	int keep = 0;
	char digs[8];
	if (up >= 0 && up < 60) {
		uint64_t f = mant << up;
		for (int k = 0; k < 8 && f; k++) {
			f = (f & 0x0fffffffffffffffull) * 10;
			const int dg = (int)(f >> 60);
			digs[k] = (char)('0' + dg);
			if (dg) keep = k + 1;
		}
	}
	if (keep == 0) *d++ = '0';
	else { memcpy(d, digs, keep); d += keep; }
	*d = 0;
	return (int)(d - out);
}

static inline void snapshot_rgb(const radk_triangle &t, int rgb[3]) {	// rad.cpp:168-178
	for (int j = 0; j < 3; j++) {
		float f = 0;
		for (int k = 0; k < 3; k++) f += t.accB[k][j];
		f /= 3;
		float lo = (255.0f < f) ? 255.0f : f;
		f = (lo < 0.0f) ? 0.0f : lo;
		if (std::isnan(f)) f = 0;
		rgb[j] = (int)f;
	}
}

static char token[64] = {0};	// the reference's static conversion buffer survives across calls

// The three "x y z nx ny nz" lines of one leaf (rad.cpp:181-219), appended to out.
static void append_geometry(const radk_triangle &t, std::string &out) {
	std::string normal;
	for (int k = 0; k < 3; k++) {
		float_to_text(t.n[k], token);
		normal += token;
		normal += (k < 2) ? " " : "\n";
	}
	for (int v = 0; v < 3; v++) {
		for (int k = 0; k < 3; k++) {
			float_to_text(t.p[v][k], token);
			out += token;
			out += ' ';
		}
		out += normal;
	}
}

void build_geometry_cache(Model &m) {
	m.geom_text.clear();
	m.geom_off.clear();
	m.geom_slot.clear();
	m.geom_text.reserve((size_t)m.trinum * 60);
	for (int i = 0; i < m.trinum; i++) {
		const radk_triangle &t = m.tri[i];
		if (t.parent < 0) continue;
		m.geom_off.push_back(m.geom_text.size());
		m.geom_slot.push_back(i);
		append_geometry(t, m.geom_text);
	}
	m.geom_off.push_back(m.geom_text.size());
	m.geom_valid = true;
}

static inline char *put_int(char *d, int v) {	// "%d"
	if (v < 0) { *d++ = '-'; v = -v; }
	char tmp[12];
	int n = 0;
	do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
	while (n) *d++ = tmp[--n];
	return d;
}

void format_fin_colors(const Model &m, const int *rgb, std::string &out) {
	const size_t leaves = m.geom_slot.size();
	out.resize(m.geom_text.size() + leaves * 64);
	char *d = &out[0];
	for (size_t k = 0; k < leaves; k++) {
		memcpy(d, "Triangle \n", 10), d += 10;
		const int *c = rgb + (size_t)m.geom_slot[k] * 3;
		d = put_int(d, c[0]), *d++ = ' ';
		d = put_int(d, c[1]), *d++ = ' ';
		d = put_int(d, c[2]);
		memcpy(d, " 0 0 0\n", 7), d += 7;
		const size_t len = m.geom_off[k + 1] - m.geom_off[k];
		memcpy(d, m.geom_text.data() + m.geom_off[k], len), d += len;
	}
	out.resize((size_t)(d - &out[0]));
}

void format_fin(const Model &m, std::string &out) {
	out.clear();
	out.reserve((size_t)m.trinum * 120);
	char line[160];
	for (int i = 0; i < m.trinum; i++) {
		const radk_triangle &t = m.tri[i];
		if (t.parent < 0) continue;
		out += "Triangle \n";
		int rgb[3];
		snapshot_rgb(t, rgb);
		int n = snprintf(line, sizeof(line), "%d %d %d 0 0 0\n", rgb[0], rgb[1], rgb[2]);
		out.append(line, n);
		append_geometry(t, out);
	}
}

}  // namespace radk
