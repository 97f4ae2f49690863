// Fast path of the single-source shooting loop (sm_100a).
//
// Observation: the form factors of a shot depend only on WHICH patch shoots (pure geometry on a
// frozen mesh), not on the radiosity state; only the cheap per-receiver update is serial from shot
// to shot.  And the next shooters are predictable: they are (almost always) the patches with the
// largest unshot energy now.  So the loop is split into rounds of three kernels:
//
//   predict_kernel   top-K candidates by unshot energy (exact top-1 + per-group maxima), form-factor
//                    cache rows assigned to the ones not cached yet, their shooter records built (K2).
//   ff_batch_kernel  K3a for all new rows at once: (row, half, unique vertex) items; facing culls first,
//                    survivors compacted through shared memory so every lane of a warp walks a ray;
//                    K x 2U rays per launch fill the 148 SMs and hide the walk's dependent-load latency.
//   shots_kernel     persistent cooperative kernel, one CTA per SM slot: up to `budget` shots, each =
//                    K3b update from the cached row + block argmax + ONE grid barrier + every CTA
//                    reducing the per-CTA partials redundantly (K1).  It returns when the next shooter
//                    has no cached row (misprediction), on convergence, or when the budget is spent.
//
// Results are bit-identical to the per-shot kernels of shoot_kernels.cuh: same arithmetic, same
// selection rule; prediction only decides which form factors are ready, never which patch shoots.
#pragma once
#include <cooperative_groups.h>

#include "shoot_kernels.cuh"

namespace radk {

constexpr int kPredictGroups = 2048;	// groups for the approximate top-K (= 2 keys per thread of the last CTA)
constexpr int kMaxRows = 32;		// upper bound of form-factor cache rows (one ballot word)

struct ShotRec {		// the shooter of the next shot, handed from shot to shot / kernel to kernel
	int slot;		// TriStore slot, -1 none
	int row;		// cache row holding its form factors, -1 = not cached
	float dB[3];		// its unshot energy at selection time (what CalRadiosity multiplies by)
	int valid;
	int pad[2];
};

struct alignas(32) Partial {	// one CTA's best candidate of a shot
	unsigned long long key;
	float dB[3];
	int pad;
};

// ---- multi-GPU mailboxes (one PeerBox per rank, in that rank's HBM, mapped into every peer) ---------
// Ranks exchange (a) per round: their local top candidates with full shooter records, (b) per shot:
// their local best (key, deltaB).  Writers store the payload, __threadfence_system(), then a sequence
// number (run epoch << 40 | round or shot number); readers spin on the sequence number in their OWN
// box (local memory), so a wait costs one NVLink store latency.  Payload slots are double-buffered by
// parity: a writer can be at most one round / one shot ahead of the slowest reader.
constexpr int kMaxWorld = 8;
struct PredCand {
	unsigned long long key;
	ShooterRec rec;
	int pad[2];
};
// Per-shot mail is self-validating: two 16-byte halves, each carrying the shot tag and written with one
// 128-bit store, so no system-scope fence and no separate flag are needed (a half is one NVLink write).
struct alignas(32) ShotMail {
	uint4 h0;	// key lo, key hi, dB0 bits, tag
	uint4 h1;	// dB1 bits, dB2 bits, tag, 0
};
__device__ __forceinline__ void send_mail(ShotMail *dst, const Partial &p, unsigned tag) {
	const unsigned long long key = p.key;
	asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(&dst->h0), "r"((unsigned)key), "r"((unsigned)(key >> 32)),
		     "r"(__float_as_uint(p.dB[0])), "r"(tag) : "memory");
	asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(&dst->h1), "r"(__float_as_uint(p.dB[1])),
		     "r"(__float_as_uint(p.dB[2])), "r"(tag), "r"(0u) : "memory");
}
__device__ __forceinline__ bool read_mail(const ShotMail *src, unsigned tag, Partial *out) {
	uint4 a, b;
	asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(&src->h0) : "memory");
	asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(&src->h1) : "memory");
	if (a.w != tag || b.z != tag) return false;
	out->key = ((unsigned long long)a.y << 32) | a.x;
	out->dB[0] = __uint_as_float(a.z), out->dB[1] = __uint_as_float(b.x), out->dB[2] = __uint_as_float(b.y), out->pad = 0;
	return true;
}

struct PeerBox {
	ShotMail shot[2][kMaxWorld];
	PredCand pred[2][kMaxWorld][kMaxRows];
	unsigned long long pred_seq[2][kMaxWorld];
};

struct FastState {
	PeerBox *box[kMaxWorld];	// box[r] = rank r's mailbox (peer-mapped); box[rank] is local
	int rank, world;
	unsigned long long epoch;	// run epoch (bumped by reset / rewind on every rank alike)
	unsigned long long *pred_round;	// rounds predicted so far in this epoch (device)
	int *error;			// set when a peer wait timed out
	unsigned long long *dbg;	// optional [16] %globaltimer stamps of predict_kernel's phases (nullptr = off)
	float2 *ff;			// [rows][U] cached form factors: .x toward half 0, .y toward half 1
	int rows;			// cache capacity (<= kMaxRows)
	int *row_slot;			// [rows] slot whose form factors the row holds, -1 free
	int *slot_row;			// [trinum] row of a slot, -1 none
	ShooterRec *rec;		// [rows] shooter record per row
	int *todo;			// [rows] rows to fill in the next ff_batch launch
	int *ntodo;
	unsigned long long *group_key;	// [kPredictGroups]
	unsigned *ticket;
	ShotRec *handoff;		// current shooter between kernels
	Partial *partial;		// [2][max CTAs], 32-byte records
	unsigned *barrier;		// grid barrier counter (zeroed by the last CTA of predict_kernel before each shots launch)
	long long *budget;		// shots the current radk_ctx_step call may still run
	unsigned *row_mask;		// [rows][mask_words] bit n set <=> leaf n's list contains the row's logsrc
	int mask_words;
	int *stalled;			// 1: shots_kernel stopped on a cache miss
};

// Spin until *p == expected (p is in local memory, written by a peer).  Gives up after ~20 s of SM
// clock so that a dead peer cannot hang the GPU; the caller then flags an error.
__device__ __forceinline__ bool peer_wait(const unsigned long long *p, unsigned long long expected) {
	const volatile unsigned long long *vp = (const volatile unsigned long long *)p;
	const long long t0 = clock64();
	unsigned spins = 0;
	while (*vp != expected) {
		if ((++spins & 0xfff) == 0 && clock64() - t0 > 40000000000ll) return false;
	}
	__threadfence_system();
	return true;
}
__device__ __forceinline__ void peer_copy(void *dst, const void *src, int bytes) {	// 8-byte granules, volatile stores
	volatile unsigned long long *d = (volatile unsigned long long *)dst;
	const unsigned long long *q = (const unsigned long long *)src;
	for (int i = 0; i < bytes / 8; i++) d[i] = q[i];
}

// Warp 0 of one CTA: rebuild the form-factor cache around the candidate keys top[0..rows) (descending,
// 0 = none): rows whose slot is no longer a candidate are freed, candidates without a row get the free
// rows in order; F.todo / *F.ntodo list the rows to fill.  Returns (in every lane) the number of new rows;
// lane j's row is written to *my_row.
__device__ int rebuild_cache(const FastState &F, const unsigned long long *top, int lane, int *my_row) {
	const int rows = F.rows;
	const unsigned long long key = lane < rows ? top[lane] : 0ull;
	const bool valid = key != 0ull;
	const int slot = (int)(key & 0xffffffffu);
	int row = valid ? F.slot_row[slot] : -1;
	unsigned keepmask = 0u;
	for (int j = 0; j < 32; j++) {
		const int rj = __shfl_sync(0xffffffffu, row, j);
		if (rj >= 0) keepmask |= 1u << rj;
	}
	int held = lane < rows ? F.row_slot[lane] : 0;		// slot held by row `lane`
	if (lane < rows && held >= 0 && !((keepmask >> lane) & 1u)) {
		F.slot_row[held] = -1, F.row_slot[lane] = -1, held = -1;
	}
	const unsigned freemask = __ballot_sync(0xffffffffu, lane < rows && held < 0);
	const bool need = valid && row < 0;
	const unsigned needmask = __ballot_sync(0xffffffffu, need);
	const int rank = __popc(needmask & ((1u << lane) - 1u));
	const int nfree = __popc(freemask), nneed = __popc(needmask);
	if (need && rank < nfree) {
		row = (int)__fns(freemask, 0, rank + 1);
		F.row_slot[row] = slot, F.slot_row[slot] = row, F.todo[rank] = row;
	}
	const int nt = nneed < nfree ? nneed : nfree;
	if (lane == 0) *F.ntodo = nt;
	*my_row = row;
	return nt;
}

// Leaf masks of the new rows: which leaves list the shooter's original triangle.  One warp per row, all rows
// in flight together (the chain todo -> record -> inverted index -> leaves is four dependent loads).
__device__ void build_row_masks(const DevScene &S, const FastState &F, int nt) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	for (int t = warp; t < nt; t += nwarps) {
		const int row = F.todo[t];
		unsigned *mask = F.row_mask + (size_t)row * F.mask_words;
		for (int w = lane; w < F.mask_words; w += 32) mask[w] = 0u;
		__syncwarp();
		const int logsrc = F.rec[row].logsrc;
		const int b = S.tree.tri_leaf_begin[logsrc], e = S.tree.tri_leaf_begin[logsrc + 1];
		for (int k = b + lane; k < e; k += 32) {
			const int leaf = S.tree.tri_leaves[k];
			atomicOr(&mask[leaf >> 5], 1u << (leaf & 31));
		}
	}
}

// ---- predict ------------------------------------------------------------------------------------
__device__ __forceinline__ void dbg_stamp(const FastState &F, int i) {
	if (F.dbg && threadIdx.x == 0) {
		unsigned long long t;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
		F.dbg[i] = t;
	}
}

__global__ void __launch_bounds__(1024) predict_kernel(DevScene S, FastState F) {
	__shared__ bool s_last;
	__shared__ int s_ntodo;
	if (blockIdx.x == 0) dbg_stamp(F, 0);
	const int T = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
	// thread t only sees receivers r == t (mod T); T is a multiple of kPredictGroups, so its running
	// maximum belongs to group t % kPredictGroups
	unsigned long long best = 0ull;
	for (int r = gtid; r < S.R; r += T) {
		unsigned c;
		const unsigned long long k = candidate_key(S, S.st[0][r].w, S.st[1][r].w, S.st[2][r].w, __ldg(&S.slot[r]), ~0ull, &c);
		best = k > best ? k : best;
	}
	if (best) atomicMax(&F.group_key[gtid & (kPredictGroups - 1)], best);
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) s_last = (atomicAdd(F.ticket, 1u) == gridDim.x - 1);
	__syncthreads();
	if (!s_last) return;
	__threadfence();
	if (threadIdx.x == 0) *F.barrier = 0u;	// the round's shots_kernel counts its grid barriers from zero
	dbg_stamp(F, 1);
	// ---- last CTA: the `rows` largest group maxima.  Threads 0..255 own eight keys each; per round a thread
	// maximum, an 8-warp reduction and one named barrier of 256 threads - the other 24 warps wait below.
	__shared__ unsigned long long wmax[2][8];
	__shared__ unsigned long long top[kMaxRows];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int rows = F.rows;
	if (threadIdx.x < kMaxRows) top[threadIdx.x] = 0ull;
	__syncthreads();
	if (threadIdx.x < 256) {
		unsigned long long k[kPredictGroups / 256];
#pragma unroll
		for (int q = 0; q < kPredictGroups / 256; q++) {
			k[q] = __ldcg(&F.group_key[threadIdx.x + 256 * q]);
			F.group_key[threadIdx.x + 256 * q] = 0ull;	// ready for the next round
		}
		for (int it = 0; it < rows; it++) {
			unsigned long long mine = 0ull;
#pragma unroll
			for (int q = 0; q < kPredictGroups / 256; q++) mine = k[q] > mine ? k[q] : mine;
			const unsigned long long wm = warp_max_u64(mine);
			if (lane == 0) wmax[it & 1][warp] = wm;
			asm volatile("bar.sync 1, 256;" ::: "memory");
			unsigned long long m = 0ull;
#pragma unroll
			for (int w = 0; w < 8; w++) m = wmax[it & 1][w] > m ? wmax[it & 1][w] : m;
			if (m == 0ull) break;		// fewer candidates than rows (uniform across the 256 threads)
			if (mine == m) {
#pragma unroll
				for (int q = 0; q < kPredictGroups / 256; q++)
					if (k[q] == m) k[q] = 0ull;
			}
			if (threadIdx.x == 0) top[it] = m;
		}
	}
	__syncthreads();
	dbg_stamp(F, 2);
	if (F.world > 1) {
		// multi-GPU: publish the local candidates (key + complete shooter record) to every rank; the cache is
		// rebuilt by predict_merge_kernel from the union
		if (threadIdx.x == 0) *F.ticket = 0u;
		__shared__ unsigned long long s_round;
		if (threadIdx.x == 0) s_round = ++(*F.pred_round);
		__syncthreads();
		const int par = (int)(s_round & 1ull);
		if (threadIdx.x < rows) {
			PredCand pc;
			pc.key = top[threadIdx.x], pc.pad[0] = pc.pad[1] = 0;
			pc.rec.valid = 0;
			// the record only travels when no rank has it yet (the caches are identical on all ranks)
			const bool fresh = pc.key != 0ull && F.slot_row[(int)(pc.key & 0xffffffffu)] < 0;
			if (fresh) build_shooter(S, S.slot_to_recv[(int)(pc.key & 0xffffffffu)], &pc.rec);
			for (int t = 0; t < F.world; t++) {
				PredCand *dst = &F.box[t]->pred[par][F.rank][threadIdx.x];
				if (fresh) {
					peer_copy(dst, &pc, (int)sizeof(PredCand));
				} else {	// key + current unshot energy (the handoff needs it if this candidate is the global best)
					*((volatile unsigned long long *)&dst->key) = pc.key;
					if (pc.key != 0ull) {
						const int r = S.slot_to_recv[(int)(pc.key & 0xffffffffu)];
						*((volatile float *)&dst->rec.dB[0]) = S.st[0][r].w;
						*((volatile float *)&dst->rec.dB[1]) = S.st[1][r].w;
						*((volatile float *)&dst->rec.dB[2]) = S.st[2][r].w;
					}
				}
			}
		}
		__syncthreads();
		if (threadIdx.x == 0) __threadfence_system();	// cumulative: orders the CTA's mail before the sequence numbers
		__syncthreads();
		if (threadIdx.x < F.world)
			*((volatile unsigned long long *)&F.box[threadIdx.x]->pred_seq[par][F.rank]) = (F.epoch << 40) | s_round;
		return;
	}
	// ---- warp 0 rebuilds the cache around the candidates: lane j owns candidate j and row j
	if (warp == 0) {
		if (lane == 0) *F.ticket = 0u;
		int row;
		const int nt = rebuild_cache(F, top, lane, &row);
		if (lane == 0) {
			s_ntodo = nt;
			// the exact argmax (top[0]) is the next shooter
			ShotRec h;
			h.pad[0] = h.pad[1] = 0;
			if (top[0] == 0ull) {
				h.slot = -1, h.row = -1, h.dB[0] = h.dB[1] = h.dB[2] = 0.0f, h.valid = 0;
				*S.sel.converged = 1;
			} else {
				h.slot = (int)(top[0] & 0xffffffffu);
				h.row = row;
				const int r = S.slot_to_recv[h.slot];
				h.dB[0] = S.st[0][r].w, h.dB[1] = S.st[1][r].w, h.dB[2] = S.st[2][r].w;
				h.valid = 1;
			}
			*F.handoff = h;
			*F.stalled = 0;
		}
	}
	__syncthreads();
	dbg_stamp(F, 3);
	const int nt = s_ntodo;
	if (threadIdx.x < nt) {	// K2 for the new rows, one thread each
		const int row = F.todo[threadIdx.x];
		build_shooter(S, S.slot_to_recv[F.row_slot[row]], &F.rec[row]);
	}
	__syncthreads();
	dbg_stamp(F, 4);
	build_row_masks(S, F, nt);
	__syncthreads();
	dbg_stamp(F, 5);
}

// Multi-GPU second half of the prediction (one CTA of 256 threads): wait for every rank's candidates,
// take the global top `rows`, rebuild the cache (identically on every rank), adopt the shooter records.
__global__ void __launch_bounds__(256) predict_merge_kernel(DevScene S, FastState F) {
	__shared__ unsigned long long wmax[2][8];
	__shared__ unsigned long long top[kMaxRows];
	__shared__ int top_src[kMaxRows];
	__shared__ int s_ntodo, s_ok;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned long long round = *F.pred_round;
	const int par = (int)(round & 1ull);
	PeerBox *mine = F.box[F.rank];
	if (threadIdx.x == 0) s_ok = 1;
	__syncthreads();
	if (threadIdx.x < F.world && !peer_wait(&mine->pred_seq[par][threadIdx.x], (F.epoch << 40) | round)) s_ok = 0;
	__syncthreads();
	if (!s_ok) {
		if (threadIdx.x == 0) {
			*F.error = 1, *F.ntodo = 0;
			ShotRec h = {-1, -1, {0.f, 0.f, 0.f}, 0, {0, 0}};
			*F.handoff = h;
		}
		return;
	}
	const int rows = F.rows;
	// thread t owns candidate (src = t / 32, j = t % 32)
	const int src = threadIdx.x >> 5, j = threadIdx.x & 31;
	unsigned long long k0 = 0ull;
	if (src < F.world && j < rows) k0 = *((volatile unsigned long long *)&mine->pred[par][src][j].key);
	for (int it = 0; it < rows; it++) {
		const unsigned long long wm = warp_max_u64(k0);
		if (lane == 0) wmax[it & 1][warp] = wm;
		__syncthreads();
		const unsigned long long m = warp_max_u64(lane < 8 ? wmax[it & 1][lane] : 0ull);
		if (m != 0ull && k0 == m) k0 = 0ull, top_src[it] = threadIdx.x;
		if (threadIdx.x == 0) top[it] = m;
	}
	__syncthreads();
	if (warp == 0) {
		int row;
		const int nt = rebuild_cache(F, top, lane, &row);
		if (lane == 0) {
			s_ntodo = nt;
			ShotRec h;
			h.pad[0] = h.pad[1] = 0;
			if (top[0] == 0ull) {
				h.slot = -1, h.row = -1, h.dB[0] = h.dB[1] = h.dB[2] = 0.0f, h.valid = 0;
				*S.sel.converged = 1;
			} else {
				const PredCand *pc = &mine->pred[par][top_src[0] >> 5][top_src[0] & 31];
				h.slot = (int)(top[0] & 0xffffffffu);
				h.row = row;
				h.dB[0] = *((volatile float *)&pc->rec.dB[0]), h.dB[1] = *((volatile float *)&pc->rec.dB[1]),
				h.dB[2] = *((volatile float *)&pc->rec.dB[2]);
				h.valid = 1;
			}
			*F.handoff = h;
			*F.stalled = 0;
		}
	}
	__syncthreads();
	const int nt = s_ntodo;
	// adopt the shooter records of the new rows from the mailbox (one warp per row, 4-byte granules)
	for (int t = warp; t < nt; t += 8) {
		const int row = F.todo[t];
		const int slot = F.row_slot[row];
		int which = -1;
		for (int i = 0; i < rows; i++)
			if ((int)(top[i] & 0xffffffffu) == slot && top[i] != 0ull) which = top_src[i];
		const volatile int *from = (const volatile int *)&mine->pred[par][which >> 5][which & 31].rec;
		int *to = (int *)&F.rec[row];
		for (int w = lane; w < (int)(sizeof(ShooterRec) / 4); w += 32) to[w] = from[w];
	}
	__syncthreads();
	if (threadIdx.x < nt) F.rec[F.todo[threadIdx.x]].recv = S.slot_to_recv[F.rec[F.todo[threadIdx.x]].slot];
	__syncthreads();
	build_row_masks(S, F, nt);
}

// ---- batched form factors -----------------------------------------------------------------------
constexpr int kFFBlock = 128;	// small CTAs: 128 x 2 vertices (measured 64 / 128 / 256 / 512 threads: church 3.37 / 3.19 / 3.23 / 3.51 ms)
constexpr int kFFItems = 2;	// candidate vertices per thread before compaction (measured on church / box / room: 1: 3.94 / 1.38 / 3.72 ms,
				// 2: 3.59 / 1.27 / 3.61, 4: 3.61 / 1.32 / 3.69, 8: 3.92 / 1.47 / 4.07)

// 48 warps per SM (40 registers): measured 32 / 40 / 48 / 64 warps per SM on church 4.06 / 3.46 / 3.20 / 3.35 ms (64 = 32
// registers, spills)
__global__ void __launch_bounds__(kFFBlock, 1536 / kFFBlock) ff_batch_kernel(DevScene S, FastState F) {
	const int j = blockIdx.z;
	if (j >= *F.ntodo) return;
	const int row = F.todo[j], half = blockIdx.y;
	__shared__ ShooterRec sh;
	__shared__ int q_u[kFFBlock * kFFItems];
	__shared__ int wcount[kFFItems * (kFFBlock / 32)];
	__shared__ int s_next;
	if (threadIdx.x < sizeof(ShooterRec) / 4) ((int *)&sh)[threadIdx.x] = ((const int *)(F.rec + row))[threadIdx.x];
	const unsigned *row_mask = F.row_mask + (size_t)row * F.mask_words;	// bit n: leaf n lists the shooter's original (read through L1)
	if (threadIdx.x == 0) s_next = 0;
	__syncthreads();
	float2 *out = F.ff + (size_t)row * S.U;
	const int base = blockIdx.x * (kFFBlock * kFFItems);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (!sh.splittable && half == 1) {	// the single centre sample is handled by half 0
		for (int it = 0; it < kFFItems; it++) {
			const int u = base + it * kFFBlock + threadIdx.x;
			if (u < S.U) out[u].y = 0.0f;
		}
		return;
	}
	const FFSample *smp = sh.splittable ? &sh.half[half] : &sh.centre;	// stays in shared memory (registers)
	float *outh = (float *)out + half;	// outh[2 * u] = form factor of vertex u toward this half
	// phase 1: culls + analytic form factor; survivors are queued in vertex order.  The form factor itself is
	// stored right away - it is the value if the ray reaches the shooter; a blocked ray zeroes it in phase 2.
	bool live[kFFItems];
	unsigned ball[kFFItems];
#pragma unroll
	for (int it = 0; it < kFFItems; it++) {
		const int u = base + it * kFFBlock + threadIdx.x;
		live[it] = false;
		if (u < S.U) {
			const float4 v = __ldg(&S.uv[u]);
			const float4 pn = __ldg(&S.par_n[__float_as_int(v.w)]);
			float dx, dy, dz;
			const float ff = ff_geom(*smp, pn.x, pn.y, pn.z, pn.w, v.x, v.y, v.z, dx, dy, dz, live[it]);
			outh[2 * (size_t)u] = live[it] ? ff : 0.0f;
		}
		ball[it] = __ballot_sync(0xffffffffu, live[it]);
		if (lane == 0) wcount[it * (kFFBlock / 32) + warp] = __popc(ball[it]);
	}
	__syncthreads();
	int total = 0;
#pragma unroll
	for (int it = 0; it < kFFItems; it++) {
		int off = 0;
		for (int w = 0; w < kFFItems * (kFFBlock / 32); w++) {
			const int c = wcount[w];
			if (w < it * (kFFBlock / 32) + warp) off += c;
			if (it == 0) total += c;
		}
		if (live[it]) q_u[off + __popc(ball[it] & ((1u << lane) - 1u))] = base + it * kFFBlock + threadIdx.x;
	}
	__syncthreads();
	// phase 2: every lane walks rays from the queue; a lane whose ray is decided fetches the next one at the
	// following leaf step.  One leaf step per iteration, in sub-phases that each end with a warp-wide reconvergence
	// (lanes that leave a leaf list early wait for the scans of the others instead of running NextPoint on their
	// own: ncu showed it at 11 of 32 lanes); leaf_visit() runs leaves with and without the shooter's original
	// through the same bounded any-hit scan (the separate exact scan cost 9.6 % of the instructions at 5.9 lanes).
	// The per-lane state is kept small (the probe point is recomputed from t, counters are packed, the sample
	// point and the result live in shared / global memory): registers decide how many CTAs an SM holds, and
	// this latency-bound walk gains from every extra warp.
	const DevTree &T = S.tree;
	int u = -1, ban = 0, prev = -1;	// u < 0: no ray
	unsigned sg = 0;	// leaf lookups so far | "precision exceeded" retries << 11
	float px = 0, py = 0, pz = 0, vx = 0, vy = 0, vz = 0, t = 0;
	bool more = true;
	for (;;) {
		if (u < 0 && more) {
			const int i = atomicAdd(&s_next, 1);
			if (i >= total) {
				more = false;
			} else {
				u = q_u[i];
				const float4 v = __ldg(&S.uv[u]);
				px = v.x, py = v.y, pz = v.z, ban = __float_as_int(v.w);
				vx = smp->cx - px, vy = smp->cy - py, vz = smp->cz - pz;
				prev = -1, sg = 0, t = 0.f;
			}
		}
		if (!__any_sync(0xffffffffu, u >= 0)) break;
		int verdict = -1;	// -1 undecided, 0 blocked / left the scene, 1 shooter visible
		int cur = -1;
		bool visit = false;
		if (u >= 0) {	// phase A: leaf lookup
			if ((sg & 0x7ffu) == 0) cur = __ldg(&S.uv_leaf[u]);	// the first probe point is the vertex itself: its leaf is static
			else cur = leaf_of<false>(T, px + t * vx, py + t * vy, pz + t * vz, nullptr);
			sg++;
			if (cur < 0 || (sg & 0x7ffu) > (unsigned)kMaxWalkSteps) {
				verdict = 0;
			} else if (prev == cur) {		// raycast.cpp:344-351
				t = (float)((double)t + 1.0);
				sg += 1u << 11;
				if ((sg >> 11) > 100000u) verdict = 0;
			} else {
				visit = true;
			}
		}
		__syncwarp();
		if (visit)	// phase B: the leaf's list
			verdict = leaf_visit(T, cur, (__ldg(&row_mask[cur >> 5]) >> (cur & 31)) & 1u, px, py, pz, vx, vy, vz, ban, sh.logsrc);
		__syncwarp();
		if (visit && verdict < 0) {	// phase C: NextPoint (raycast.cpp:235-251)
			const float4 lo = leaf_lo(T, cur);
			const float4 hi = leaf_hi(T, cur);
			prev = cur;
			float tn = 1000000.0f, s;
			if (fabsf(vx) > 1e-6f) { s = ((vx > 0 ? hi.x : lo.x) - px) / vx; tn = (s < tn) ? s : tn; }
			if (fabsf(vy) > 1e-6f) { s = ((vy > 0 ? hi.y : lo.y) - py) / vy; tn = (s < tn) ? s : tn; }
			if (fabsf(vz) > 1e-6f) { s = ((vz > 0 ? hi.z : lo.z) - pz) / vz; tn = (s < tn) ? s : tn; }
			t = (float)((double)tn + 0.1);
		}
		if (verdict >= 0) {	// phase D: a visible ray keeps the form factor phase 1 stored
			if (verdict == 0) outh[2 * (size_t)u] = 0.0f;
			u = -1;
		}
	}
}

// ---- persistent shot loop -----------------------------------------------------------------------
// Arrival with release semantics (publishes what the calling thread wrote before, i.e. its CTA's partial record) and
// waiting with acquire loads: no separate MEMBAR.GPU on either side (church: 4.8 -> 4.5 us per shot).
__device__ __forceinline__ void barrier_arrive(unsigned *counter) {
	asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void barrier_wait(const unsigned *counter, unsigned target) {
	unsigned v;
	do {
		asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
	} while (v < target);
}
constexpr int kShotsBlock = 512;

// CalRadiosity (shade.cpp:244-259) of one receiver from cached vertex form factors.
__device__ __forceinline__ void apply_shot(float4 &a0, float4 &a1, float4 &a2, const int4 &id, const float4 &fk,
					   const float2 *__restrict__ ffrow, float sdB0, float sdB1, float sdB2) {
	const float2 h0 = __ldcg(&ffrow[id.x]), h1 = __ldcg(&ffrow[id.y]), h2 = __ldcg(&ffrow[id.z]);
	const float ff[3] = {h0.x + h0.y, h1.x + h1.y, h2.x + h2.y};
	const float lo[3] = {fk.x * sdB0 / 3, fk.y * sdB1 / 3, fk.z * sdB2 / 3};	// shade.cpp:247-249
	float dB[3] = {a0.w, a1.w, a2.w};
	float4 *acc[3] = {&a0, &a1, &a2};
#pragma unroll
	for (int v = 0; v < 3; v++) {
		float d;
		d = lo[0] * ff[v]; dB[0] += d; acc[v]->x += d;
		d = lo[1] * ff[v]; dB[1] += d; acc[v]->y += d;
		d = lo[2] * ff[v]; dB[2] += d; acc[v]->z += d;
	}
	a0.w = dB[0], a1.w = dB[1], a2.w = dB[2];
}
__device__ __forceinline__ bool unsaturated(const float4 &a0, const float4 &a1, const float4 &a2) {	// rad.cpp:424-428
	return a0.x < 255.0f || a0.y < 255.0f || a0.z < 255.0f || a1.x < 255.0f || a1.y < 255.0f || a1.z < 255.0f ||
	       a2.x < 255.0f || a2.y < 255.0f || a2.z < 255.0f;
}

// RPT > 0: every thread keeps the state, vertex ids and reflectance of RPT receivers in registers for
// the whole launch (the shard fits gridDim * 512 * RPT); RPT == 0: receivers are streamed from L2/HBM
// every shot (any shard size).
template <int RPT>
__global__ void __launch_bounds__(kShotsBlock, RPT == 1 ? 2 : 1) shots_kernel(DevScene S, FastState F, int max_ctas) {
	__shared__ ShotRec cur;
	__shared__ unsigned long long s_key[kShotsBlock / 32];
	__shared__ float s_dB[kShotsBlock / 32][3];
	__shared__ Partial s_best;
	__shared__ int s_row_slot[kMaxRows];			// cache directory: slot held by each row
	constexpr int NR = RPT > 0 ? RPT : 1;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int nctas = gridDim.x;
	// contiguous chunk of receivers per CTA
	const int per = (S.R + nctas - 1) / nctas;
	const int r_begin = blockIdx.x * per, r_end = min(S.R, r_begin + per);
	if (threadIdx.x == 0) {
		cur = *F.handoff;
		if (cur.valid) cur.row = F.slot_row[cur.slot];
	}
	if (threadIdx.x < kMaxRows) s_row_slot[threadIdx.x] = threadIdx.x < F.rows ? F.row_slot[threadIdx.x] : -1;
	// resident receivers
	float4 ra0[NR], ra1[NR], ra2[NR], rfk[NR];
	int4 rid[NR];
	int rslot[NR];
	if (RPT > 0) {
#pragma unroll
		for (int k = 0; k < NR; k++) {
			const int r = r_begin + threadIdx.x + k * kShotsBlock;
			rslot[k] = -1;
			if (r < r_end) {
				ra0[k] = S.st[0][r], ra1[k] = S.st[1][r], ra2[k] = S.st[2][r];
				rid[k] = __ldg(&S.vid[r]);
				rfk[k] = __ldg(&S.par_fk[rid[k].w]);
				rslot[k] = __ldg(&S.slot[r]);
			}
		}
	}
	__syncthreads();
	long long budget = *F.budget;
	const long long shots_before = *S.sel.shots_done;
	unsigned epoch = 0;
	int parity = 0;
	long long done = 0;
	int stalled = 0;
	unsigned long long tA = 0, tB = 0, tC = 0, tD = 0, tmark = 0;
	const bool dbg = F.dbg && blockIdx.x == 0 && threadIdx.x == 0;
	auto now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
	while (true) {
		if (!cur.valid || budget <= 0) break;
		if (cur.row < 0) { stalled = 1; break; }
		if (dbg) tmark = now();
		const float2 *ffrow = F.ff + (size_t)cur.row * S.U;
		const int shooter_slot = cur.slot;
		const float sdB0 = cur.dB[0], sdB1 = cur.dB[1], sdB2 = cur.dB[2];
		unsigned long long key = 0ull;
		float kd0 = 0.f, kd1 = 0.f, kd2 = 0.f;
		if (RPT > 0) {
#pragma unroll
			for (int k = 0; k < NR; k++) {
				if (rslot[k] < 0) continue;
				if (rslot[k] == shooter_slot) ra0[k].w = ra1[k].w = ra2[k].w = 0.0f;	// rad.cpp:446-449
				else if (unsaturated(ra0[k], ra1[k], ra2[k])) apply_shot(ra0[k], ra1[k], ra2[k], rid[k], rfk[k], ffrow, sdB0, sdB1, sdB2);
				unsigned c;
				const unsigned long long kk = candidate_key(S, ra0[k].w, ra1[k].w, ra2[k].w, rslot[k], ~0ull, &c);
				if (kk > key) key = kk, kd0 = ra0[k].w, kd1 = ra1[k].w, kd2 = ra2[k].w;
			}
		} else {
			// streamed receivers (HBM bound: 96 B of state + 20 B of ids per receiver and shot): kStreamUnroll receivers
			// per thread are in flight together - all their loads are issued before the first update, the vertex
			// ids and form-factor gathers of one overlap the state loads of the others
			constexpr int kStreamUnroll = 4;
			for (int r0 = r_begin + threadIdx.x; r0 < r_end; r0 += kShotsBlock * kStreamUnroll) {
				float4 a0[kStreamUnroll], a1[kStreamUnroll], a2[kStreamUnroll], fk[kStreamUnroll];
				int4 id[kStreamUnroll];
				int slot[kStreamUnroll];
#pragma unroll
				for (int k = 0; k < kStreamUnroll; k++) {
					const int r = r0 + k * kShotsBlock;
					slot[k] = -1;
					if (r < r_end) {
						a0[k] = S.st[0][r], a1[k] = S.st[1][r], a2[k] = S.st[2][r];
						slot[k] = __ldg(&S.slot[r]);
						id[k] = __ldg(&S.vid[r]);
					}
				}
#pragma unroll
				for (int k = 0; k < kStreamUnroll; k++)
					if (slot[k] >= 0) fk[k] = __ldg(&S.par_fk[id[k].w]);
#pragma unroll
				for (int k = 0; k < kStreamUnroll; k++) {
					const int r = r0 + k * kShotsBlock;
					if (slot[k] < 0) continue;
					if (slot[k] == shooter_slot) {
						a0[k].w = a1[k].w = a2[k].w = 0.0f;
						S.st[0][r].w = 0.0f, S.st[1][r].w = 0.0f, S.st[2][r].w = 0.0f;
					} else if (unsaturated(a0[k], a1[k], a2[k])) {
						apply_shot(a0[k], a1[k], a2[k], id[k], fk[k], ffrow, sdB0, sdB1, sdB2);
						S.st[0][r] = a0[k], S.st[1][r] = a1[k], S.st[2][r] = a2[k];
					}
					unsigned c;
					const unsigned long long kk = candidate_key(S, a0[k].w, a1[k].w, a2[k].w, slot[k], ~0ull, &c);
					if (kk > key) key = kk, kd0 = a0[k].w, kd1 = a1[k].w, kd2 = a2[k].w;
				}
			}
		}
		if (dbg) { const unsigned long long t = now(); tA += t - tmark; tmark = t; }
		// CTA argmax carrying the winner's unshot energy
		const unsigned long long wmax = warp_max_u64(key);
		const unsigned src = __ffs(__ballot_sync(0xffffffffu, key == wmax && wmax != 0ull)) - 1;
		if (wmax != 0ull) {
			kd0 = __shfl_sync(0xffffffffu, kd0, src), kd1 = __shfl_sync(0xffffffffu, kd1, src), kd2 = __shfl_sync(0xffffffffu, kd2, src);
		}
		if (lane == 0) s_key[warp] = wmax, s_dB[warp][0] = kd0, s_dB[warp][1] = kd1, s_dB[warp][2] = kd2;
		__syncthreads();
		if (warp == 0) {
			unsigned long long k = lane < kShotsBlock / 32 ? s_key[lane] : 0ull;
			const unsigned long long m = warp_max_u64(k);
			const unsigned w = __ffs(__ballot_sync(0xffffffffu, k == m)) - 1;	// some lane < 16 holds m (possibly 0)
			if (lane == 0) {
				Partial p;
				p.key = m, p.dB[0] = s_dB[w][0], p.dB[1] = s_dB[w][1], p.dB[2] = s_dB[w][2], p.pad = 0;
				s_best = p;
				F.partial[parity * max_ctas + blockIdx.x] = p;
				// the CTA's record is final here - arrive at the (local) grid barrier right away
				barrier_arrive(F.barrier);
			}
		}
		__syncthreads();	// s_key / s_dB are reused below
		epoch++;
		if (F.world == 1) {
			// single GPU: publish the CTA's best, one grid barrier, then every CTA reduces all records the same way
			// (identical `cur` everywhere, no second barrier).  The whole 32-byte record comes with one pair of
			// 128-bit loads, and the winner's cache row is looked up in the shared-memory directory, so the
			// barrier is followed by a single L2 round trip.
			if (dbg) { const unsigned long long t = now(); tB += t - tmark; tmark = t; }
			if (threadIdx.x == 0) barrier_wait(F.barrier, epoch * (unsigned)nctas);
			__syncthreads();
			if (dbg) { const unsigned long long t = now(); tC += t - tmark; tmark = t; }
			unsigned long long k = 0ull;
			float d0 = 0.f, d1 = 0.f, d2 = 0.f;
			if (threadIdx.x < nctas) {
				const uint4 *rec = (const uint4 *)&F.partial[parity * max_ctas + threadIdx.x];
				const uint4 a = __ldcg(rec), bb = __ldcg(rec + 1);
				k = ((unsigned long long)a.y << 32) | a.x;
				d0 = __uint_as_float(a.z), d1 = __uint_as_float(a.w), d2 = __uint_as_float(bb.x);
			}
			const unsigned long long m1 = warp_max_u64(k);
			if (lane == 0) s_key[warp] = m1;
			__syncthreads();
			const unsigned long long m = warp_max_u64(lane < kShotsBlock / 32 ? s_key[lane] : 0ull);	// every warp reduces the warp maxima
			// the next shooter is written in place: slot and unshot energy by the thread that holds the winning record,
			// its cache row (shared-memory directory) by warp 0 - one barrier ends the shot
			if (m != 0ull && k == m) cur.slot = (int)(m & 0xffffffffu), cur.dB[0] = d0, cur.dB[1] = d1, cur.dB[2] = d2, cur.valid = 1;
			if (warp == 0) {
				const unsigned hit = __ballot_sync(0xffffffffu, m != 0ull && s_row_slot[lane] == (int)(m & 0xffffffffu));
				if (lane == 0) {
					if (blockIdx.x == 0) S.sel.shot_log[(shots_before + done) % S.sel.shot_log_cap] = shooter_slot;
					cur.row = hit ? __ffs(hit) - 1 : -1;
					if (m == 0ull) cur.slot = -1, cur.row = -1, cur.dB[0] = cur.dB[1] = cur.dB[2] = 0.0f, cur.valid = 0;
				}
			}
			__syncthreads();
		} else {
			// multi-GPU: CTA 0 waits for the local CTAs, reduces their partials and mails the rank's best to every
			// rank; every CTA (on every rank) then waits for all ranks' mail in its own box.
			const unsigned tag = (unsigned)F.epoch * 0x9E3779B1u + (unsigned)(shots_before + done + 1);	// never equals a stale slot's tag
			if (blockIdx.x == 0) {
				if (threadIdx.x == 0) barrier_wait(F.barrier, epoch * (unsigned)nctas);
				__syncthreads();
				unsigned long long k = 0ull;
				int who = -1;
				for (int b = threadIdx.x; b < nctas; b += kShotsBlock) {
					const unsigned long long kb = __ldcg(&F.partial[parity * max_ctas + b].key);
					if (kb > k) k = kb, who = b;
				}
				const unsigned long long m1 = warp_max_u64(k);
				if (lane == 0) s_key[warp] = m1;
				__syncthreads();
				unsigned long long m = 0ull;
				for (int w = 0; w < kShotsBlock / 32; w++) m = s_key[w] > m ? s_key[w] : m;
				if (threadIdx.x == 0) s_best.key = 0ull, s_best.dB[0] = s_best.dB[1] = s_best.dB[2] = 0.0f, s_best.pad = 0;
				__syncthreads();
				if (m != 0ull && k == m) {
					const Partial *p = &F.partial[parity * max_ctas + who];
					s_best.key = m, s_best.dB[0] = __ldcg(&p->dB[0]), s_best.dB[1] = __ldcg(&p->dB[1]), s_best.dB[2] = __ldcg(&p->dB[2]);
				}
				__syncthreads();
				if (threadIdx.x < F.world) send_mail(&F.box[threadIdx.x]->shot[parity][F.rank], s_best, tag);
			}
			__shared__ int s_ok;
			if (threadIdx.x == 0) s_ok = 1;
			__syncthreads();
			PeerBox *mine = F.box[F.rank];
			__shared__ Partial s_mail[kMaxWorld];
			if (threadIdx.x < F.world) {
				const long long t0 = clock64();
				unsigned spins = 0;
				while (!read_mail(&mine->shot[parity][threadIdx.x], tag, &s_mail[threadIdx.x])) {
					if ((++spins & 0xfff) == 0 && clock64() - t0 > 40000000000ll) { s_ok = 0; break; }
				}
			}
			__syncthreads();
			if (threadIdx.x == 0) {
				Partial best;
				best.key = 0ull, best.dB[0] = best.dB[1] = best.dB[2] = 0.0f, best.pad = 0;
				if (s_ok) {
					for (int t = 0; t < F.world; t++)
						if (s_mail[t].key > best.key) best = s_mail[t];
				} else {
					*F.error = 1;
				}
				s_best = best;
			}
			__syncthreads();
			if (threadIdx.x == 0) {
				if (blockIdx.x == 0) S.sel.shot_log[(shots_before + done) % S.sel.shot_log_cap] = shooter_slot;
				ShotRec n;
				n.pad[0] = n.pad[1] = 0;
				if (s_best.key == 0ull) {
					n.slot = -1, n.row = -1, n.dB[0] = n.dB[1] = n.dB[2] = 0.0f, n.valid = 0;
				} else {
					n.slot = (int)(s_best.key & 0xffffffffu);
					n.row = -1;
					for (int j = 0; j < F.rows; j++)
						if (s_row_slot[j] == n.slot) n.row = j;
					n.dB[0] = s_best.dB[0], n.dB[1] = s_best.dB[1], n.dB[2] = s_best.dB[2];
					n.valid = 1;
				}
				cur = n;
			}
			__syncthreads();
		}
		if (dbg) { const unsigned long long t = now(); tD += t - tmark; tmark = t; }
		done++, budget--, parity ^= 1;
	}
	if (RPT > 0) {	// write the resident state back
#pragma unroll
		for (int k = 0; k < NR; k++) {
			const int r = r_begin + threadIdx.x + k * kShotsBlock;
			if (rslot[k] >= 0) S.st[0][r] = ra0[k], S.st[1][r] = ra1[k], S.st[2][r] = ra2[k];
		}
	}
	if (dbg) { F.dbg[8] += tA, F.dbg[9] += tB, F.dbg[10] += tC, F.dbg[11] += tD, F.dbg[12] += (unsigned long long)done; }
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		*F.handoff = cur;
		*F.budget = budget;
		*S.sel.shots_done += done;
		atomicAdd(&S.stats->shots, (unsigned long long)done);
		*F.stalled = stalled;
		if (!cur.valid) *S.sel.converged = 1;
	}
}

}  // namespace radk
