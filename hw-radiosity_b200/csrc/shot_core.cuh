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

// x / 3 (the `/ 3` of shade.cpp:247-249), correctly rounded without the division: q0 = x * RN(1/3), the residual
// r = x - 3 q0 exactly (one FMA), q = RN(q0 + r * RN(1/3)).  Checked against x / 3.0f for ALL 2^32 float bit patterns
// (the CPU test tier samples it, tests/hostemu; the exhaustive run takes 18 s): identical except -0 and +-inf, which are passed
// through.  The IEEE division costs ~35 instructions on the device, three of them per receiver and shot (and per
// candidate and shot in the chain's dependent loop).
RADK_HD float div3(float x) {
	const float y = 0.3333333432674407958984375f;	// RN(1/3) = 0x3eaaaaab
	const float q0 = x * y;
#if defined(__CUDA_ARCH__)
	const float r = __fmaf_rn(-3.0f, q0, x);
	const float q = __fmaf_rn(r, y, q0);
	return (x == 0.0f || fabsf(x) == __int_as_float(0x7f800000)) ? x : q;
#else
	const float r = fmaf(-3.0f, q0, x);
	const float q = fmaf(r, y, q0);
	return (x == 0.0f || fabsf(x) == INFINITY) ? x : q;
#endif
}

// CalRadiosity (shade.cpp:244-259) of one receiver: a_v = accB[v].rgb | deltaB[v]; ff_v = vertex form factors;
// fk = Frgb * (0.5/255); sdB = the shooter's unshot energy.
RADK_HD void apply_shot_ff(float4 &a0, float4 &a1, float4 &a2, float fkx, float fky, float fkz, float ff0, float ff1, float ff2,
			   float sdB0, float sdB1, float sdB2) {
	const float ff[3] = {ff0, ff1, ff2};
	const float lo[3] = {div3(fkx * sdB0), div3(fky * sdB1), div3(fkz * sdB2)};	// shade.cpp:247-249: fk * dB / 3
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

// One scheduled shot: who shoots, from which cache row, with which unshot energy.  With -parallel_src P > 1 an
// iteration of the reference's loop (rad.cpp:406-451) has up to P co-shooters: its entries are consecutive, the first
// carries kSchedHead, the last kSchedLast (P = 1: every entry carries both).  next_key (on the last entry of an
// iteration) is the smallest key among the shooters scheduled for the NEXT iteration (0: none scheduled) - a receiver
// whose own key exceeds it after this iteration would have been selected instead and breaks the schedule.
constexpr int kSchedHead = 1, kSchedLast = 2;
struct alignas(32) ShotSched {
	int slot, row;
	float dB[3];
	int flags;
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

// listCanditate's shooter count (rad.cpp:400): n = min((|C| + 4) / 5, ParallelSource), |C| = receivers whose unshot
// energy exceeds the limit.  The chain knows |C| exactly only at the start of a round (the selection pass counts it);
// afterwards it knows a lower bound (each shooter leaves C, receivers may enter it).  n is certain while the bound
// gives P already: lb >= 5 P - 4.
RADK_HD int shooters_for(long long count, int P) {
	const long long n = (count + 4) / 5;
	return (int)(n < (long long)P ? n : (long long)P);
}
RADK_HD bool shooter_count_certain(long long count_lb, int P) { return P <= 1 || count_lb >= 5ll * P - 4; }

// Sequential restatement of the chain (the device runs it with one warp, fast_kernels.cuh): replays the iterations on
// the candidate records only.  `count` = |C| at the start of the round (only read when P > 1), max_iters = iterations
// the caller may still run.  Returns the number of schedule entries (<= kMaxChain); *iters_out = iterations scheduled.
inline int chain_schedule_host(CandRec *c, int rows, float converge, long long max_iters, int P, long long count, ShotSched *out,
			       int *iters_out = nullptr) {
	unsigned long long key[kMaxRows];
	int sel[kMaxRows];
	auto keys = [&]() {
		for (int i = 0; i < rows; i++) key[i] = c[i].slot >= 0 ? energy_key(converge, c[i].a0.w, c[i].a1.w, c[i].a2.w, c[i].slot) : 0ull;
	};
	// the n largest keys, descending, into sel[]; false when the candidates cannot supply n shooters
	auto select = [&](int n) {
		bool used[kMaxRows] = {};
		for (int j = 0; j < n; j++) {
			int b = -1;
			for (int i = 0; i < rows; i++)
				if (!used[i] && key[i] != 0ull && (b < 0 || key[i] > key[b])) b = i;
			if (b < 0) return false;
			used[b] = true, sel[j] = b;
		}
		return true;
	};
	int k = 0, iters = 0, last = -1;
	long long lb = count;
	keys();
	while ((long long)iters < max_iters) {
		if (iters > 0 && !shooter_count_certain(lb, P)) break;
		const int n = P <= 1 ? 1 : shooters_for(lb, P);
		if (n <= 0 || k + n > kMaxChain || !select(n)) break;
		if (last >= 0) out[last].next_key = key[sel[n - 1]];
		for (int j = 0; j < n; j++) {
			const CandRec &s = c[sel[j]];
			ShotSched e;
			e.slot = s.slot, e.row = sel[j], e.dB[0] = s.a0.w, e.dB[1] = s.a1.w, e.dB[2] = s.a2.w;
			e.flags = (j == 0 ? kSchedHead : 0) | (j == n - 1 ? kSchedLast : 0);
			e.next_key = 0ull;
			out[k + j] = e;
		}
		for (int i = 0; i < rows; i++) {
			if (c[i].slot < 0) continue;
			bool shoots = false;
			for (int j = 0; j < n; j++) shoots |= sel[j] == i;
			if (shoots) continue;		// rad.cpp:421-422: co-shooters do not receive
			if (!unsaturated(c[i].a0, c[i].a1, c[i].a2)) continue;	// rad.cpp:424-428: checked once per iteration
			for (int j = 0; j < n; j++) {
				const int b = sel[j];
				apply_shot_ff(c[i].a0, c[i].a1, c[i].a2, c[i].fk[0], c[i].fk[1], c[i].fk[2], c[i].fsum[b][0], c[i].fsum[b][1],
					      c[i].fsum[b][2], out[k + j].dB[0], out[k + j].dB[1], out[k + j].dB[2]);
			}
		}
		for (int j = 0; j < n; j++) c[sel[j]].a0.w = c[sel[j]].a1.w = c[sel[j]].a2.w = 0.0f;	// rad.cpp:446-449
		last = k + n - 1;
		k += n, iters++, lb -= n;
		keys();
	}
	if (last >= 0 && out[last].next_key == 0ull) {
		// key of the shooter that would come next if the round went on (P = 1: the violation test of the last shot is
		// harmless either way - nothing scheduled follows it)
		unsigned long long best = 0ull;
		for (int i = 0; i < rows; i++) best = key[i] > best ? key[i] : best;
		out[last].next_key = P <= 1 ? best : ~0ull;
	}
	if (iters_out) *iters_out = iters;
	return k;
}

// The apply pass of one receiver: entries [0, n) of the schedule (n ends an iteration) on (a0, a1, a2).  ff(row, vertex
// id) returns the cached form factor pair of a unique vertex.  Returns the index of the LAST entry of the first
// iteration after which the schedule no longer holds for this receiver (n: none): its key exceeds next_key, or it left
// the candidate set without shooting (a NaN energy; the chain's lower bound on |C| would be wrong).
template <typename FF>
RADK_HD int apply_schedule_multi(float4 &a0, float4 &a1, float4 &a2, int id0, int id1, int id2, float fkx, float fky, float fkz,
			   int slot, const ShotSched *sched, int n, float converge, FF ff) {
	int viol = n;
	// the form factors of a shot do not depend on the receiver's state: those of kChunk shots are fetched together
	// (independent loads in flight) before the shots are applied in order
	constexpr int kChunk = 8;
	bool shoots = false, receive = false, was_cand = false;
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
			if (s.flags & kSchedHead) {
				// is this receiver one of the iteration's shooters (rad.cpp:421-422) ?  eligible otherwise (424-428) ?
				shoots = slot == s.slot;
				for (int q = k; !(sched[q].flags & kSchedLast); q++) shoots |= slot == sched[q + 1].slot;
				receive = !shoots && unsaturated(a0, a1, a2);
				was_cand = energy_key(converge, a0.w, a1.w, a2.w, slot) != 0ull;
			}
			if (receive) apply_shot_ff(a0, a1, a2, fkx, fky, fkz, f[j][0], f[j][1], f[j][2], s.dB[0], s.dB[1], s.dB[2]);
			if (s.flags & kSchedLast) {
				if (shoots) a0.w = a1.w = a2.w = 0.0f;
				const unsigned long long key = energy_key(converge, a0.w, a1.w, a2.w, slot);
				if (viol == n && (key > s.next_key || (was_cand && !shoots && key == 0ull))) viol = k;
			}
		}
	}
	return viol;
}

// Single shooter per iteration (no -parallel_src): every entry is an iteration of its own.
template <typename FF>
RADK_HD int apply_schedule_single(float4 &a0, float4 &a1, float4 &a2, int id0, int id1, int id2, float fkx, float fky, float fkz,
				  int slot, const ShotSched *sched, int n, float converge, FF ff) {
	int viol = n;
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

template <typename FF>
RADK_HD int apply_schedule(float4 &a0, float4 &a1, float4 &a2, int id0, int id1, int id2, float fkx, float fky, float fkz,
			   int slot, const ShotSched *sched, int n, float converge, FF ff, bool multi = true) {
	return multi ? apply_schedule_multi(a0, a1, a2, id0, id1, id2, fkx, fky, fkz, slot, sched, n, converge, ff)
		     : apply_schedule_single(a0, a1, a2, id0, id1, id2, fkx, fky, fkz, slot, sched, n, converge, ff);
}

}  // namespace radk
