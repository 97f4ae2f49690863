// Per-receiver part of one shot and the batched ("chained") shot schedule.
//
// The reference's loop (src/rad.cpp:499-514) is strictly serial: listCanditate (rad.cpp:387-404) picks the
// patch with the largest unshot energy, parallelRadiosity (406-451) applies CalRadiosity (shade.cpp:244-259)
// to every receiver, the shooter's deltaB is zeroed (446-449), repeat.  The energies only move through
// form factors that are already known once a patch's cache row is filled, so the order in which the cached
// candidates shoot can be worked out on a handful of records WITHOUT touching the receivers:
//
//   chain     the 32 candidates (cache rows) carry their own radiosity state and the form factors between
//             each other (a 32 x 32 x 3 table); replaying the shots on those records alone gives the schedule
//             (slot, row, deltaB of every shot) as long as no receiver outside the candidate set overtakes;
//   apply     every receiver applies the whole schedule in one pass (its gathers are independent of each other)
//             and notes the first shot after which its own key exceeds the key of the scheduled next shooter
//             - the first violation of the assumption above;
//   commit    shots up to and including the first violation (minimum over all receivers and ranks) are exactly
//             the reference's; if the schedule went further, the receivers are restored from the copy the apply
//             pass left behind and the valid prefix is applied again.
//
// The arithmetic per receiver and shot is the reference's, operation by operation, whichever pass runs it,
// so results stay bit-identical.  The functions are __host__ __device__: tests/hostemu runs the same code on
// the CPU (one process or several gloo ranks) against the oracle.
#pragma once
#include "rt_core.cuh"

namespace radk {

#ifndef RADK_MAX_ROWS
#define RADK_MAX_ROWS 128
#endif
#ifndef RADK_MAX_CHAIN
#define RADK_MAX_CHAIN 192
#endif
constexpr int kMaxRows = RADK_MAX_ROWS;	// form-factor cache rows = candidates per round (a multiple of 32: kMaxRows / 32 per lane of the chain warp)
constexpr int kMaxChain = RADK_MAX_CHAIN;	// shots scheduled per round at most (< 256: the violation mail carries it in 8 bits)
static_assert(kMaxRows % 32 == 0 && kMaxRows <= 128, "one chain warp holds kMaxRows / 32 <= 4 candidates per lane");
static_assert(kMaxChain < 255, "first violation travels in 8 bits");

// key = (order-preserving bits of delta) << 32 | slot : lexicographic max == the reference's
// sort-ascending-take-from-the-back over pair<float,int> (rad.cpp:398-402).
RADK_HD unsigned ordered_bits(float f) {
	const unsigned u = (unsigned)float_as_int_hd(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
RADK_HD unsigned long long make_key(float delta, int slot) {
	return ((unsigned long long)ordered_bits(delta) << 32) | (unsigned)slot;
}
// 0 when the receiver is no candidate (rad.cpp:392: delta > ConvergeLimit)
RADK_HD unsigned long long energy_key(float converge, float d0, float d1, float d2, int slot) {
	const float delta = d0 + d1 + d2;
	if (!(delta > converge)) return 0ull;
	return make_key(delta, slot);
}

RADK_HD bool unsaturated(const float4 &a0, const float4 &a1, const float4 &a2) {	// rad.cpp:424-428
	return a0.x < 255.0f || a0.y < 255.0f || a0.z < 255.0f || a1.x < 255.0f || a1.y < 255.0f || a1.z < 255.0f ||
	       a2.x < 255.0f || a2.y < 255.0f || a2.z < 255.0f;
}

// CalRadiosity (shade.cpp:244-259) of one receiver: a_v = accB[v].rgb | deltaB[v]; ff_v = vertex form factors;
// fk = Frgb * (0.5/255); sdB = the shooter's unshot energy.
RADK_HD void apply_shot_ff(float4 &a0, float4 &a1, float4 &a2, float fkx, float fky, float fkz, float ff0, float ff1, float ff2,
			   float sdB0, float sdB1, float sdB2) {
	const float ff[3] = {ff0, ff1, ff2};
	const float lo[3] = {fkx * sdB0 / 3, fky * sdB1 / 3, fkz * sdB2 / 3};	// shade.cpp:247-249
	float dB[3] = {a0.w, a1.w, a2.w};
	float4 *acc[3] = {&a0, &a1, &a2};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int v = 0; v < 3; v++) {
		float d;
		d = lo[0] * ff[v]; dB[0] += d; acc[v]->x += d;
		d = lo[1] * ff[v]; dB[1] += d; acc[v]->y += d;
		d = lo[2] * ff[v]; dB[2] += d; acc[v]->z += d;
	}
	a0.w = dB[0], a1.w = dB[1], a2.w = dB[2];
}

// One scheduled shot: who shoots, from which cache row, with which unshot energy, and the key of the shooter
// scheduled after it (0: none) - a receiver whose key exceeds next_key after this shot breaks the schedule.
struct alignas(32) ShotSched {
	int slot, row;
	float dB[3];
	int pad;
	unsigned long long next_key;
};

// A candidate's record as its owner publishes it: radiosity state, reflectance, slot (-1: row unused) and the form
// factors (both halves summed) of every cache row at its three vertices.
struct alignas(16) CandRec {
	float4 a0, a1, a2;
	float fk[3];
	int slot;
	float fsum[kMaxRows][3];
};

// Sequential restatement of the chain (the device runs it with one warp lane per candidate, fast_kernels.cuh):
// replays shots on the candidate records only.  Returns the number of scheduled shots (<= max_shots, kMaxChain).
inline int chain_schedule_host(CandRec *c, int rows, float converge, long long max_shots, ShotSched *out) {
	unsigned long long key[kMaxRows];
	auto best_of = [&](int *who) {
		unsigned long long best = 0ull;
		for (int i = 0; i < rows; i++) {
			key[i] = c[i].slot >= 0 ? energy_key(converge, c[i].a0.w, c[i].a1.w, c[i].a2.w, c[i].slot) : 0ull;
			if (key[i] > best) best = key[i], *who = i;
		}
		return best;
	};
	int b = -1, k = 0;
	unsigned long long best = best_of(&b);
	while (best != 0ull && k < kMaxChain && (long long)k < max_shots) {
		ShotSched s;
		s.slot = c[b].slot, s.row = b, s.dB[0] = c[b].a0.w, s.dB[1] = c[b].a1.w, s.dB[2] = c[b].a2.w, s.pad = 0;
		for (int i = 0; i < rows; i++) {
			if (c[i].slot < 0) continue;
			if (c[i].slot == s.slot) c[i].a0.w = c[i].a1.w = c[i].a2.w = 0.0f;	// rad.cpp:446-449
			else if (unsaturated(c[i].a0, c[i].a1, c[i].a2))
				apply_shot_ff(c[i].a0, c[i].a1, c[i].a2, c[i].fk[0], c[i].fk[1], c[i].fk[2], c[i].fsum[b][0], c[i].fsum[b][1],
					      c[i].fsum[b][2], s.dB[0], s.dB[1], s.dB[2]);
		}
		best = best_of(&b);
		s.next_key = best;
		out[k++] = s;
	}
	return k;
}

// The apply pass of one receiver: shots [0, n) of the schedule on (a0, a1, a2).  ff(row, vertex id) returns the cached
// form factor pair of a unique vertex.  Returns the first shot after which this receiver's key exceeds the scheduled
// next shooter's key (n: none).
template <typename FF>
RADK_HD int apply_schedule(float4 &a0, float4 &a1, float4 &a2, int id0, int id1, int id2, float fkx, float fky, float fkz,
			   int slot, const ShotSched *sched, int n, float converge, FF ff) {
	int viol = n;
	// the form factors of a shot do not depend on the receiver's state: those of kChunk shots are fetched together
	// (independent loads in flight) before the shots are applied in order
	constexpr int kChunk = 8;
	for (int k0 = 0; k0 < n; k0 += kChunk) {
		float f[kChunk][3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (int j = 0; j < kChunk; j++) {
			const int row = sched[k0 + j < n ? k0 + j : n - 1].row;
			f[j][0] = ff(row, id0), f[j][1] = ff(row, id1), f[j][2] = ff(row, id2);
		}
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (int j = 0; j < kChunk; j++) {
			const int k = k0 + j;
			if (k >= n) break;
			const ShotSched &s = sched[k];
			if (slot == s.slot) {
				a0.w = a1.w = a2.w = 0.0f;
			} else if (unsaturated(a0, a1, a2)) {
				apply_shot_ff(a0, a1, a2, fkx, fky, fkz, f[j][0], f[j][1], f[j][2], s.dB[0], s.dB[1], s.dB[2]);
			}
			if (viol == n && energy_key(converge, a0.w, a1.w, a2.w, slot) > s.next_key) viol = k;
		}
	}
	return viol;
}

}  // namespace radk
