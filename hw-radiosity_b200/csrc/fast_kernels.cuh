// Fast path of the single-source shooting loop (sm_100a).
//
// Observation: the form factors of a shot depend only on WHICH patch shoots (pure geometry on a
// frozen mesh), not on the radiosity state; only the cheap per-receiver update is serial from shot
// to shot.  And the next shooters are predictable: they are (almost always) the patches with the
// largest unshot energy now.  So the loop runs in rounds of five kernels, no host round trip in between:
//
//   predict_kernel   top-K candidates by unshot energy (exact top-1 + per-group maxima), form-factor
//                    cache rows assigned to the ones not cached yet, their shooter records built (K2).
//   ff_batch_kernel  K3a for all new rows at once: (row, half, unique vertex) items; facing culls first,
//                    survivors compacted through shared memory so every lane of a warp walks a ray.
//   chain_kernel     one CTA: collects the candidates' own records (state + the form factors between the
//                    candidates), replays the shots on those 32 records alone and writes the schedule
//                    (shot_core.cuh).  Multi-GPU: each candidate's owner publishes its record to every rank.
//   apply_kernel     K3b for the whole schedule in one pass over the receivers (state read once per round, the
//                    form-factor gathers of all shots independent of each other), leaves a copy of the old state
//                    and the index of the first shot after which a non-candidate overtakes (first violation).
//   commit_kernel    shots up to the first violation are the reference's; if the schedule went further the valid
//                    prefix is re-applied from the copy.  Bookkeeping (shot log, budget).
//
// Results are bit-identical to the per-shot kernels of shoot_kernels.cuh: same arithmetic, same
// selection rule; prediction only decides which form factors are ready and which shots can be batched,
// never which patch shoots.
#pragma once
#include <cstddef>
#include <cooperative_groups.h>

#include "shoot_kernels.cuh"

namespace radk {

// ---- TMA 1-D bulk copies (cp.async.bulk, global -> shared, completion on an mbarrier) -------------------------
// A CTA's read-only tile (the shooter record and the 4 KB of vertices it works on; the round's schedule in the
// apply pass) is fetched by ONE thread as a bulk asynchronous copy instead of one load per thread; everybody
// waits on the mbarrier's phase.  Addresses and sizes are multiples of 16 bytes.
// Measured (church / box / room, ms per run, same box): the schedule copy of the apply pass is neutral to slightly
// faster (update 0.544 -> 0.538) and is ON; staging ff_batch_kernel's tile costs 2 % (church 3.86 -> 3.94: a CTA lives
// for ~128 rays, the copy's latency sits in front of every one of its 26 000 CTAs per launch while the plain loads it
// replaces were coalesced and hit L1 on re-use) and is OFF (RADK_TMA_FF=1 builds it).
#ifndef RADK_TMA_APPLY
#define RADK_TMA_APPLY 1
#endif
#ifndef RADK_TMA_FF
#define RADK_TMA_FF 0
#endif
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned arrivals) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");	// visible to the async proxy before the copies are issued
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
		     "r"(bytes), "r"(smem_u32(bar))
		     : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" ::"r"(smem_u32(bar)),
		"r"(parity)
		: "memory");
}

constexpr int kPredictGroups = 2048;	// groups for the approximate top-K (= 2 keys per thread of the last CTA)

struct ShotRec {		// the shooter of the next shot as the prediction sees it (diagnostics)
	int slot;		// TriStore slot, -1 none
	int row;		// cache row holding its form factors, -1 = not cached
	float dB[3];		// its unshot energy at selection time (what CalRadiosity multiplies by)
	int valid;
	int pad[2];
};

// ---- multi-GPU mailboxes (one PeerBox per rank, in that rank's HBM, mapped into every peer) ---------
// Three exchanges per round, none per shot: (a) the ranks' local top candidates with full shooter records,
// (b) each candidate's record (state + form factors between the candidates) from the rank that owns it,
// (c) each rank's first violation of the round's schedule.  (a) and (b): the writer stores the payload,
// fence.sc.sys (__threadfence_system), then a sequence number (run epoch << 40 | round); readers spin on the
// sequence number in their OWN box (local memory) and fence before reading the payload - the release / acquire
// pattern of the PTX memory model.  (c) is a single 64-bit word (tag | value), one store.  Slots are
// double-buffered by round parity: a writer can be at most one round ahead of the slowest reader.
constexpr int kMaxWorld = 8;
struct PredCand {
	unsigned long long key;
	ShooterRec rec;
	int pad[2];
};
struct PeerBox {
	PredCand pred[2][kMaxWorld][kMaxRows];
	unsigned long long pred_seq[2][kMaxWorld];
	CandRec cand[2][kMaxRows];
	unsigned long long cand_seq[2][kMaxWorld];
	unsigned long long viol[2][kMaxWorld];
};

struct FastState {
	PeerBox *box[kMaxWorld];	// box[r] = rank r's mailbox (peer-mapped); box[rank] is local
	int rank, world;
	unsigned long long *pred_round;	// (run epoch << 40) | rounds predicted so far in this epoch; the epoch is bumped by reset /
					// rewind on every rank alike, so stale mail of an earlier run can never match (device memory: the
					// kernel arguments stay the same from run to run, which lets a captured graph be reused)
	unsigned long long cond;	// cudaGraphConditionalHandle of the round loop (0: rounds are launched by the host)
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
	ShotRec *handoff;		// the exact argmax as predict_kernel saw it
	long long *budget;		// shots the current radk_ctx_step call may still run
	unsigned *row_mask;		// [rows][mask_words] bit n set <=> leaf n's list contains the row's logsrc
	int mask_words;
	ShotSched *sched;		// [kMaxChain] this round's schedule (chain_kernel)
	int *nsched;			// its length
	int *viol;			// first violation found by this rank's receivers (nsched: none)
	unsigned *apply_ticket;		// last-CTA-done counter of apply_kernel (multi-GPU: the last CTA mails viol)
	unsigned *chain_ticket;		// last-CTA-done counter of chain_kernel's candidate-ray phase
	float4 *st_prev[3];		// [R] the receivers' state before the round's apply pass
	unsigned long long *counters;	// [8] 0: rows filled, 1: rounds, 2: shots scheduled, 3: shots committed, 4: rounds that re-applied
	float2 *cand_ff;		// [kMaxRows new rows][kMaxRows candidates][3 vertices] form factors of this round's NEW rows at the
					// candidates' vertices (ff_cand_kernel): lets chain_kernel run beside ff_batch_kernel
	// -parallel_src P > 1 (an iteration shoots min((|C| + 4) / 5, P) patches, rad.cpp:387-404)
	int parallel_src;
	unsigned *local_count;		// receivers of this shard above the convergence limit, counted by predict_kernel's scan
	long long *cand_count;		// |C| over all ranks at the start of the round (read by chain_kernel)
	unsigned long long *exact;	// [0] threshold T (the P-th largest group maximum), [1] list length, [2..2+kExactCap) every key >= T,
					// [2+kExactCap .. +kMaxRows) the sorted group maxima: see predict_exact_kernel
};
constexpr int kExactCap = 2048;	// keys the exact selection pass can collect (= kPredictGroups: sorted in the same shared array)

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

// All threads of one CTA (>= kMaxRows of them): rebuild the form-factor cache around the candidate keys top[0..rows)
// (descending, 0 = none; shared memory): rows whose slot is no longer a candidate are freed, candidates without a row
// get the free rows in ascending order; F.todo / *F.ntodo list the rows to fill.  Thread j < rows owns candidate j and
// row j.  Returns (in every thread) the number of new rows; s_row[j] = row of candidate j (-1 none), s_todo_cand[t] =
// candidate that todo entry t serves.  Deterministic in (top, cache directory): every rank of a sharded run rebuilds
// the same cache.
struct CacheScratch {
	unsigned keep[kMaxRows / 32];
	int fcnt[kMaxRows / 32], ncnt[kMaxRows / 32];
	int freelist[kMaxRows];
	int row[kMaxRows];
	int todo_cand[kMaxRows];
	int nt;
};
__device__ int rebuild_cache(const FastState &F, const unsigned long long *top, CacheScratch &cs) {
	const int rows = F.rows, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid < kMaxRows / 32) cs.keep[tid] = 0u;
	__syncthreads();
	const unsigned long long key = tid < rows ? top[tid] : 0ull;
	const bool valid = key != 0ull;
	const int slot = (int)(key & 0xffffffffu);
	int row = valid ? F.slot_row[slot] : -1;
	if (row >= 0) atomicOr(&cs.keep[row >> 5], 1u << (row & 31));
	__syncthreads();
	int held = tid < rows ? F.row_slot[tid] : 0;		// slot held by row `tid`
	if (tid < rows && held >= 0 && !((cs.keep[tid >> 5] >> (tid & 31)) & 1u)) {
		F.slot_row[held] = -1, F.row_slot[tid] = -1, held = -1;
	}
	const bool isfree = tid < rows && held < 0;
	const bool need = valid && row < 0;
	const unsigned fb = __ballot_sync(0xffffffffu, isfree), nb = __ballot_sync(0xffffffffu, need);
	if (lane == 0 && warp < kMaxRows / 32) cs.fcnt[warp] = __popc(fb), cs.ncnt[warp] = __popc(nb);
	__syncthreads();
	int foff = 0, noff = 0, nfree = 0, nneed = 0;
#pragma unroll
	for (int w = 0; w < kMaxRows / 32; w++) {
		if (w < warp) foff += cs.fcnt[w], noff += cs.ncnt[w];
		nfree += cs.fcnt[w], nneed += cs.ncnt[w];
	}
	if (isfree) cs.freelist[foff + __popc(fb & ((1u << lane) - 1u))] = tid;
	__syncthreads();
	if (need) {
		const int r = noff + __popc(nb & ((1u << lane) - 1u));
		if (r < nfree) {
			row = cs.freelist[r];
			F.row_slot[row] = slot, F.slot_row[slot] = row, F.todo[r] = row;
			cs.todo_cand[r] = tid;
		}
	}
	const int nt = nneed < nfree ? nneed : nfree;
	if (tid == 0) *F.ntodo = nt, cs.nt = nt;
	if (tid < kMaxRows) cs.row[tid] = row;
	__syncthreads();
	return nt;
}

// Descending bitonic sort of n = 2^k 64-bit keys in shared memory by the whole CTA (n / 2 <= blockDim.x).
__device__ void sort_keys_desc(unsigned long long *keys, int n) {
	for (int k = 2; k <= n; k <<= 1)
		for (int j = k >> 1; j > 0; j >>= 1) {
			const int t = threadIdx.x;
			if (t < n / 2) {
				const int i = (t / j) * 2 * j + (t % j), p = i + j;
				const unsigned long long a = keys[i], b = keys[p];
				const bool desc = (i & k) == 0;	// blocks alternate direction; the last merge (k = n) is all descending
				if (desc ? a < b : a > b) keys[i] = b, keys[p] = a;
			}
			__syncthreads();
		}
}

// The 64 largest of 2048 keys, sorted descending into keys[0 .. 64), by one CTA of 1024 threads (the rest of the array is
// scratch afterwards).  Each warp sorts its 64 keys in registers (bitonic network on shuffles, no block barrier), then
// five merge levels halve the number of runs: c[e] = max(a[e], b[63 - e]) is the top half of two descending runs as a
// bitonic sequence, six more compare-exchange stages sort it.  11 block barriers instead of the 66 of the full sort.
__device__ __forceinline__ void cmpx_desc(unsigned long long &x, int j, bool desc, int lane) {	// partner = lane ^ j, same register
	const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, j);
	const bool lower = (lane & j) == 0;
	x = (lower == desc) ? (x > y ? x : y) : (x < y ? x : y);
}
__device__ void topk64_desc(unsigned long long *keys) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long x0 = keys[warp * 64 + lane], x1 = keys[warp * 64 + 32 + lane];	// elements e = lane and e = lane + 32 of the warp's run
	__syncthreads();	// everybody has read its keys: the array is scratch from here on
#pragma unroll
	for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
		for (int j = k >> 1; j > 0; j >>= 1) {
			// direction of element e: descending where (e & k) == 0; e & 32 is the register index
			cmpx_desc(x0, j, k < 32 ? (lane & k) == 0 : true, lane);
			cmpx_desc(x1, j, k < 32 ? (lane & k) == 0 : false, lane);
		}
	// k = 64 (all descending): j = 32 is the register pair of a lane, then the shuffles
	if (x0 < x1) { const unsigned long long t = x0; x0 = x1, x1 = t; }
#pragma unroll
	for (int j = 16; j > 0; j >>= 1) cmpx_desc(x0, j, true, lane), cmpx_desc(x1, j, true, lane);
	for (int stride = 1; stride < 32; stride <<= 1) {
		keys[warp * 64 + lane] = x0, keys[warp * 64 + 32 + lane] = x1;
		__syncthreads();
		const bool active = (warp & (2 * stride - 1)) == 0;
		unsigned long long p0 = 0ull, p1 = 0ull;
		if (active) p0 = keys[(warp + stride) * 64 + 63 - lane], p1 = keys[(warp + stride) * 64 + 31 - lane];
		__syncthreads();	// partners' runs are read before anybody overwrites its own
		if (active) {
			x0 = x0 > p0 ? x0 : p0, x1 = x1 > p1 ? x1 : p1;	// top 64 of the union, bitonic
			if (x0 < x1) { const unsigned long long t = x0; x0 = x1, x1 = t; }
#pragma unroll
			for (int j = 16; j > 0; j >>= 1) cmpx_desc(x0, j, true, lane), cmpx_desc(x1, j, true, lane);
		}
	}
	if (warp == 0) keys[lane] = x0, keys[32 + lane] = x1;
	__syncthreads();
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
// RADK_DEBUG_PREDICT timeline: dbg[16 + 16 * (round % 64) + k] = %globaltimer when event k of the round happened
// (0 predict starts, 1 predict's last CTA ends, 2 merge starts, 3 merge ends, 4 ff_cand starts, 5 ff_batch starts,
// 6 chain starts, 7 chain ends, 8 apply starts, 9 commit starts, 10 commit's block 0 ends); dbg[15] counts rounds.
__device__ __forceinline__ void dbg_trace(const FastState &F, int k) {
	if (F.dbg) {
		unsigned long long t;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
		F.dbg[16 + 16 * (int)(F.dbg[15] & 63ull) + k] = t;
	}
}
__device__ __forceinline__ void dbg_stamp(const FastState &F, int i) {
	if (F.dbg && threadIdx.x == 0) {
		unsigned long long t;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
		F.dbg[i] = t;
	}
}

// Tail of the prediction, run by ONE CTA of 1024 threads with the candidate keys top[0 .. rows) sorted descending in
// shared memory: multi-GPU - mail the local candidates; single GPU - rebuild the cache, shooter records, leaf masks.
__device__ void predict_finish(const DevScene &S, const FastState &F, const unsigned long long *top, CacheScratch &cs) {
	const int rows = F.rows;
	dbg_stamp(F, 2);
	__shared__ unsigned long long s_round;
	if (threadIdx.x == 0) s_round = ++(*F.pred_round);	// rounds are numbered from 1 within a run epoch
	__syncthreads();
	if (F.world > 1) {
		// multi-GPU: publish the local candidates (key + complete shooter record) to every rank; the cache is
		// rebuilt by predict_merge_kernel from the union
		if (threadIdx.x == 0) *F.ticket = 0u;
		const int par = (int)(s_round & 1ull);
		if (threadIdx.x < rows) {
			PredCand pc;
			pc.key = top[threadIdx.x], pc.pad[0] = rows;	// pad[0]: every rank must run the same cache size
			pc.pad[1] = (int)*F.local_count;	// pad[1] (read from entry 0): this rank's part of |C|
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
					*((volatile int *)&dst->pad[0]) = rows;
					*((volatile int *)&dst->pad[1]) = pc.pad[1];
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
			*((volatile unsigned long long *)&F.box[threadIdx.x]->pred_seq[par][F.rank]) = s_round;
		__syncthreads();	// every thread has read the count it mails
		if (threadIdx.x == 0) *F.local_count = 0u;
		return;
	}
	// ---- rebuild the cache around the candidates: thread j owns candidate j and row j
	if (threadIdx.x == 0) *F.ticket = 0u;
	const int nt = rebuild_cache(F, top, cs);
	if (threadIdx.x == 0) {
		// the exact argmax (top[0]) is the next shooter
		ShotRec h;
		h.pad[0] = h.pad[1] = 0;
		if (top[0] == 0ull) {
			h.slot = -1, h.row = -1, h.dB[0] = h.dB[1] = h.dB[2] = 0.0f, h.valid = 0;
			*S.sel.converged = 1;
		} else {
			h.slot = (int)(top[0] & 0xffffffffu);
			h.row = cs.row[0];
			const int r = S.slot_to_recv[h.slot];
			h.dB[0] = S.st[0][r].w, h.dB[1] = S.st[1][r].w, h.dB[2] = S.st[2][r].w;
			h.valid = 1;
		}
		*F.handoff = h;
		F.counters[0] += (unsigned long long)nt, F.counters[1]++;
		*F.cand_count = (long long)*F.local_count;
		*F.local_count = 0u;
	}
	dbg_stamp(F, 3);
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

__global__ void __launch_bounds__(1024) predict_kernel(DevScene S, FastState F) {
	__shared__ bool s_last;
	if (blockIdx.x == 0) dbg_stamp(F, 0);
	if (blockIdx.x == 0 && threadIdx.x == 0) dbg_trace(F, 0);
	const int T = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
	// thread t only sees receivers r == t (mod T); T is a multiple of kPredictGroups, so its running
	// maximum belongs to group t % kPredictGroups
	unsigned long long best = 0ull;
	unsigned ncand = 0u;
	for (int r = gtid; r < S.R; r += T) {
		unsigned c;
		const unsigned long long k = candidate_key(S, S.st[0][r].w, S.st[1][r].w, S.st[2][r].w, __ldg(&S.slot[r]), ~0ull, &c);
		best = k > best ? k : best;
		ncand += c;
	}
	if (best) atomicMax(&F.group_key[gtid & (kPredictGroups - 1)], best);
	if (F.parallel_src > 1) {	// |C| of listCanditate (rad.cpp:400)
		ncand = __reduce_add_sync(0xffffffffu, ncand);
		if ((threadIdx.x & 31) == 0 && ncand) atomicAdd(F.local_count, ncand);
	}
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) s_last = (atomicAdd(F.ticket, 1u) == gridDim.x - 1);
	__syncthreads();
	if (!s_last) return;
	__threadfence();
	dbg_stamp(F, 1);
	// ---- last CTA: the `rows` largest group maxima = the head of a descending sort of the group keys
	__shared__ unsigned long long keys[kPredictGroups];
	__shared__ CacheScratch cs;
	for (int q = threadIdx.x; q < kPredictGroups; q += blockDim.x) {
		keys[q] = __ldcg(&F.group_key[q]);
		F.group_key[q] = 0ull;	// ready for the next round
	}
	__syncthreads();
	if (F.parallel_src <= 1 && F.rows <= 64 && kPredictGroups == 2048) {
		if (F.dbg) {	// RADK_DEBUG_PREDICT: self-check of the partial selection against the full sort, mismatches in dbg[14]
			__shared__ unsigned long long full[kPredictGroups];
			for (int q = threadIdx.x; q < kPredictGroups; q += blockDim.x) full[q] = keys[q];
			__syncthreads();
			sort_keys_desc(full, kPredictGroups);
			topk64_desc(keys);
			if (threadIdx.x < 64 && full[threadIdx.x] != keys[threadIdx.x]) atomicAdd(&F.dbg[14], 1ull);
			__syncthreads();
		} else {
			topk64_desc(keys);
		}
	} else {
		sort_keys_desc(keys, kPredictGroups);
	}
	if (F.parallel_src > 1) {
		// group maxima can hide the 2nd .. P-th largest key behind a larger one of the same group: the exact pass
		// (predict_exact_kernel) collects every key >= T = the P-th largest group maximum and finishes the round
		if (threadIdx.x == 0) {
			F.exact[0] = keys[F.parallel_src - 1];
			F.exact[1] = 0ull;
			*F.ticket = 0u;
		}
		if (threadIdx.x < kMaxRows) F.exact[2 + kExactCap + threadIdx.x] = keys[threadIdx.x];
		return;
	}
	predict_finish(S, F, keys, cs);
	if (threadIdx.x == 0) dbg_trace(F, 1);
}

// -parallel_src P > 1, second pass of the selection: every key >= T (the P-th largest group maximum; T = 0: every
// candidate) is collected; the true P largest keys are among them.  The last CTA sorts the collected keys, appends the
// group maxima below T (they only serve as speculative candidates) and finishes the round like predict_kernel does.
__global__ void __launch_bounds__(1024) predict_exact_kernel(DevScene S, FastState F) {
	__shared__ bool s_last;
	const int T = gridDim.x * blockDim.x;
	const unsigned long long thr = F.exact[0];
	unsigned long long *list = F.exact + 2;
	const int lane = threadIdx.x & 31;
	for (int r0 = blockIdx.x * blockDim.x; r0 < S.R; r0 += T) {	// uniform trip count per CTA: warp votes inside
		const int r = r0 + threadIdx.x;
		unsigned long long k = 0ull;
		if (r < S.R) {
			unsigned c;
			k = candidate_key(S, S.st[0][r].w, S.st[1][r].w, S.st[2][r].w, __ldg(&S.slot[r]), ~0ull, &c);
		}
		const bool take = k != 0ull && k >= thr;
		const unsigned ball = __ballot_sync(0xffffffffu, take);
		if (ball) {
			unsigned long long pos = 0ull;
			if (lane == 0) pos = atomicAdd(&F.exact[1], (unsigned long long)__popc(ball));
			pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(ball & ((1u << lane) - 1u));
			if (take && pos < (unsigned long long)kExactCap) list[pos] = k;
		}
	}
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) s_last = (atomicAdd(F.ticket, 1u) == gridDim.x - 1);
	__syncthreads();
	if (!s_last) return;
	__threadfence();
	__shared__ unsigned long long keys[kExactCap];
	__shared__ CacheScratch cs;
	const unsigned long long m64 = __ldcg(&F.exact[1]);
	if (m64 > (unsigned long long)kExactCap) {
		// more keys above the threshold than the pass can sort (thousands of receivers hiding behind < P group maxima):
		// fail loudly instead of selecting inexactly
		if (threadIdx.x == 0) *F.error = 2, *F.ntodo = 0, *F.ticket = 0u, *F.local_count = 0u;
		return;
	}
	const int m = (int)m64;
	for (int q = threadIdx.x; q < kExactCap; q += blockDim.x) keys[q] = q < m ? __ldcg(&list[q]) : 0ull;
	__syncthreads();
	sort_keys_desc(keys, kExactCap);
	// group maxima below the threshold follow (the sorted group maxima are descending; those >= T are already in the list)
	if (threadIdx.x == 0 && thr != 0ull) {
		int n = m;
		const unsigned long long *gm = F.exact + 2 + kExactCap;
		for (int q = 0; q < kMaxRows && n < kMaxRows; q++) {
			const unsigned long long g = gm[q];
			if (g != 0ull && g < thr) keys[n++] = g;
		}
	}
	__syncthreads();
	predict_finish(S, F, keys, cs);
}

// Multi-GPU second half of the prediction (one CTA of 1024 threads): wait for every rank's candidates,
// take the global top `rows`, rebuild the cache (identically on every rank), adopt the shooter records.
__global__ void __launch_bounds__(1024) predict_merge_kernel(DevScene S, FastState F) {
	static_assert(kMaxWorld * kMaxRows == 1024, "one mailed candidate per thread");
	__shared__ unsigned long long keys[2 * kMaxWorld * kMaxRows];	// upper half: zero padding for the 2048-key partial selection
	__shared__ int top_src[kMaxRows];
	__shared__ CacheScratch cs;
	__shared__ int s_ok;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned long long round = *F.pred_round;
	const int par = (int)(round & 1ull);
	PeerBox *mine = F.box[F.rank];
	if (threadIdx.x == 0) s_ok = 1, dbg_trace(F, 2);
	__syncthreads();
	if (threadIdx.x < F.world) {
		if (!peer_wait(&mine->pred_seq[par][threadIdx.x], round)) s_ok = 0;
		// every rank must rebuild the same cache: a rank that runs a different number of rows is an error
		else if (*((volatile int *)&mine->pred[par][threadIdx.x][0].pad[0]) != F.rows) s_ok = 0;
	}
	__syncthreads();
	if (!s_ok) {
		if (threadIdx.x == 0) {
			*F.error = 11, *F.ntodo = 0;	// 11: the candidates of a rank did not arrive (or its cache size differs)
			ShotRec h = {-1, -1, {0.f, 0.f, 0.f}, 0, {0, 0}};
			*F.handoff = h;
		}
		return;
	}
	const int rows = F.rows;
	// thread t owns candidate (src = t / kMaxRows, j = t % kMaxRows); a slot lives on one rank, so non-zero keys are unique
	const int src = threadIdx.x / kMaxRows, j = threadIdx.x % kMaxRows;
	unsigned long long k0 = 0ull;
	if (src < F.world && j < rows) k0 = *((volatile unsigned long long *)&mine->pred[par][src][j].key);
	keys[threadIdx.x] = k0, keys[kMaxWorld * kMaxRows + threadIdx.x] = 0ull;
	if (threadIdx.x < kMaxRows) top_src[threadIdx.x] = 0;
	__syncthreads();
	// up to 64 rows: the 64 largest in registers (topk64_desc) instead of the full sort of the 1024 mailed keys
	const bool partial = rows <= 64;
	if (partial) topk64_desc(keys);
	else sort_keys_desc(keys, kMaxWorld * kMaxRows);
	const unsigned long long *top = keys;
	if (k0 != 0ull) {	// where did my key end up?  (binary search in the descending part)
		int lo = 0, hi = (partial ? 64 : kMaxWorld * kMaxRows) - 1;
		while (lo < hi) {
			const int mid = (lo + hi) >> 1;
			if (top[mid] > k0) lo = mid + 1; else hi = mid;
		}
		if (lo < rows && top[lo] == k0) top_src[lo] = threadIdx.x;
	}
	__syncthreads();
	const int nt = rebuild_cache(F, top, cs);
	if (threadIdx.x == 0) {
		ShotRec h;
		h.pad[0] = h.pad[1] = 0;
		if (top[0] == 0ull) {
			h.slot = -1, h.row = -1, h.dB[0] = h.dB[1] = h.dB[2] = 0.0f, h.valid = 0;
			*S.sel.converged = 1;
		} else {
			const PredCand *pc = &mine->pred[par][top_src[0] / kMaxRows][top_src[0] % kMaxRows];
			h.slot = (int)(top[0] & 0xffffffffu);
			h.row = cs.row[0];
			h.dB[0] = *((volatile float *)&pc->rec.dB[0]), h.dB[1] = *((volatile float *)&pc->rec.dB[1]),
			h.dB[2] = *((volatile float *)&pc->rec.dB[2]);
			h.valid = 1;
		}
		*F.handoff = h;
		F.counters[0] += (unsigned long long)nt, F.counters[1]++;
		long long count = 0;
		for (int t = 0; t < F.world; t++) count += (long long)*((volatile int *)&mine->pred[par][t][0].pad[1]);
		*F.cand_count = count;
	}
	// adopt the shooter records of the new rows from the mailbox (one warp per row, 4-byte granules)
	for (int t = warp; t < nt; t += (int)(blockDim.x >> 5)) {
		const int row = F.todo[t];
		const int which = top_src[cs.todo_cand[t]];
		const volatile int *from = (const volatile int *)&mine->pred[par][which / kMaxRows][which % kMaxRows].rec;
		int *to = (int *)&F.rec[row];
		for (int w = lane; w < (int)(sizeof(ShooterRec) / 4); w += 32) to[w] = from[w];
	}
	__syncthreads();
	if (threadIdx.x < nt) F.rec[F.todo[threadIdx.x]].recv = S.slot_to_recv[F.rec[F.todo[threadIdx.x]].slot];
	__syncthreads();
	build_row_masks(S, F, nt);
	if (threadIdx.x == 0) dbg_trace(F, 3);
}

// ---- batched form factors -----------------------------------------------------------------------
#ifndef RADK_FF_BLOCK
#define RADK_FF_BLOCK 128
#endif
// Threshold refill (round 2): idle lanes fetch new rays only when fewer than 16 lanes of the warp still walk, 4 candidate
// vertices per thread.  Measured against per-lane refill with 2 per thread (ms per run, same box): church 3.85 -> 3.78,
// box 2.07 -> 2.01, room 3.76 -> 3.67; threshold 8: 3.80 / 2.00 / 3.65, 24: 3.84 / 2.05 / 3.81; 8 per thread: 4.11 / 2.17 / 3.95.
#ifndef RADK_FF_REFILL_T
#define RADK_FF_REFILL_T 16
#endif
#ifndef RADK_FF_ITEMS
#define RADK_FF_ITEMS 4
#endif
constexpr int kFFBlock = RADK_FF_BLOCK;	// small CTAs: 128 x 2 vertices (measured 64 / 128 / 256 / 512 threads: church 3.37 / 3.19 / 3.23 / 3.51 ms)
constexpr int kFFItemsMax = RADK_FF_ITEMS;	// candidate vertices per thread before compaction (measured on church / box / room: 1: 3.94 / 1.38 / 3.72 ms,
				// 2: 3.59 / 1.27 / 3.61, 4: 3.61 / 1.32 / 3.69, 8: 3.92 / 1.47 / 4.07)

// 48 warps per SM (40 registers): measured 32 / 40 / 48 / 64 warps per SM on church 4.06 / 3.46 / 3.20 / 3.35 ms (64 = 32
// registers, spills)
// ITEMS = candidate vertices per thread (tile = kFFBlock * ITEMS vertices): 4 where the launch still has several waves of
// CTAs, fewer on small shards (ff_items_for below)
template <int ITEMS>
__global__ void __launch_bounds__(kFFBlock, 1536 / kFFBlock) ff_batch_kernel(DevScene S, FastState F) {
	constexpr int kFFItems = ITEMS;
	const int j = blockIdx.z;
	if (blockIdx.x == 0 && blockIdx.y == 0 && j == 0 && threadIdx.x == 0) dbg_trace(F, 5);
	if (j >= *F.ntodo) return;
	const int row = F.todo[j], half = blockIdx.y;
	__shared__ __align__(16) ShooterRec sh;
	__shared__ int q_u[kFFBlock * kFFItems];
	__shared__ int wcount[kFFItems * (kFFBlock / 32)];
	__shared__ int s_next;
	const int base = blockIdx.x * (kFFBlock * kFFItems);
#if RADK_TMA_FF
	// the CTA's tile - its shooter record and its vertices - by two bulk copies of one thread
	static_assert(sizeof(ShooterRec) % 16 == 0, "bulk copies move multiples of 16 bytes");
	__shared__ __align__(16) float4 s_uv[kFFBlock * kFFItems];
	__shared__ __align__(8) unsigned long long s_bar;
	if (threadIdx.x == 0) {
		s_next = 0;
		mbar_init(&s_bar, 1);
		const int nv = min(kFFBlock * kFFItems, S.U - base);
		mbar_expect_tx(&s_bar, (unsigned)(sizeof(ShooterRec) + nv * sizeof(float4)));
		tma_load_1d(&sh, F.rec + row, (unsigned)sizeof(ShooterRec), &s_bar);
		tma_load_1d(s_uv, S.uv + base, (unsigned)(nv * sizeof(float4)), &s_bar);
	}
	__syncthreads();	// the barrier is initialised before anybody polls it
	mbar_wait(&s_bar, 0u);
#else
	if (threadIdx.x < sizeof(ShooterRec) / 4) ((int *)&sh)[threadIdx.x] = ((const int *)(F.rec + row))[threadIdx.x];
	if (threadIdx.x == 0) s_next = 0;
	__syncthreads();
#endif
	const unsigned *row_mask = F.row_mask + (size_t)row * F.mask_words;	// bit n: leaf n lists the shooter's original (read through L1)
	float2 *out = F.ff + (size_t)row * S.U;
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
#if RADK_TMA_FF
			const float4 v = s_uv[it * kFFBlock + threadIdx.x];
#else
			const float4 v = __ldg(&S.uv[u]);
#endif
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
#if RADK_FF_REFILL_T < 32
		// idle lanes fetch new rays only when fewer than RADK_FF_REFILL_T lanes of the warp still walk: the fetched rays
		// start their walks together instead of each lane drifting into its own phase
		const bool refill = __popc(__ballot_sync(0xffffffffu, u >= 0)) < RADK_FF_REFILL_T;
		if (u < 0 && more && refill) {
#else
		if (u < 0 && more) {
#endif
			const int i = atomicAdd(&s_next, 1);
			if (i >= total) {
				more = false;
			} else {
				u = q_u[i];
#if RADK_TMA_FF
				const float4 v = s_uv[u - base];
#else
				const float4 v = __ldg(&S.uv[u]);
#endif
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

// Vertices per thread for a launch over U vertices x 2 halves x rows: the largest of 4 / 2 / 1 that still gives four waves
// of CTAs (12 CTAs per SM x 148 SMs); small shards (church on 8 GPUs: 13 k vertices per rank) otherwise end in a
// nearly empty last wave.
inline int ff_items_for(int U, int rows) {
	for (int items = kFFItemsMax; items > 1; items >>= 1) {
		const long long tiles = (U + kFFBlock * items - 1) / (kFFBlock * items);
		if (tiles * 2 * rows >= 4ll * 12 * 148) return items;
	}
	return 1;
}
inline const void *ff_batch_fn(int items) {
	return items >= 4 ? (const void *)ff_batch_kernel<4> : (items == 2 ? (const void *)ff_batch_kernel<2> : (const void *)ff_batch_kernel<1>);
}

// ---- form factors of the new rows at the candidates' own vertices ---------------------------------------
// chain_kernel only needs the form factors BETWEEN the candidates (a rows x rows x 3 table).  Those of rows cached
// in earlier rounds are in the cache already; those of the rows this round fills are computed here, ray by ray
// (at most rows x rows x 3 x 2 rays, the same arithmetic as ff_batch_kernel: ff_geom + the bounded any-hit walk),
// so that the chain runs concurrently with ff_batch_kernel instead of after it.
// Work item q of todo entry t = (candidate i, vertex v, half).
__device__ void ff_cand_item(const DevScene &S, const FastState &F, int t, int q) {
	const int i = q / 6, v = (q % 6) >> 1, half = q & 1;
	const int cslot = F.row_slot[i];
	const int r = cslot >= 0 ? S.slot_to_recv[cslot] : -1;
	if (r < 0) return;	// free row, or a candidate another rank owns
	const int row = F.todo[t];
	const ShooterRec &sh = F.rec[row];
	const int4 id = __ldg(&S.vid[r]);
	const int u = v == 0 ? id.x : (v == 1 ? id.y : id.z);
	float ff = 0.0f;
	if (sh.splittable || half == 0) {
		const FFSample &smp = sh.splittable ? sh.half[half] : sh.centre;
		const float4 p = __ldg(&S.uv[u]);
		const float4 pn = __ldg(&S.par_n[__float_as_int(p.w)]);
		float dx, dy, dz;
		bool cast;
		const float g = ff_geom(smp, pn.x, pn.y, pn.z, pn.w, p.x, p.y, p.z, dx, dy, dz, cast);
		if (cast && ray_sees(S.tree, F.row_mask + (size_t)row * F.mask_words, p.x, p.y, p.z, dx, dy, dz, __float_as_int(p.w), sh.logsrc)) ff = g;
	}
	((float *)&F.cand_ff[((size_t)t * kMaxRows + i) * 3 + v])[half] = ff;
}

// ---- chained shot loop ----------------------------------------------------------------------------
// word of the per-round violation mail: run epoch (16 bits) | round (40 bits) | first violation (8 bits)
__device__ __forceinline__ unsigned long long viol_word(unsigned long long round, int v) {	// round = epoch << 40 | round number
	return (round << 8) | (unsigned long long)(v & 0xff);
}
__device__ __forceinline__ float ff_sum(const float2 *ff, int U, int row, int u) {
	const float2 h = __ldg(&ff[(size_t)row * U + u]);
	return h.x + h.y;
}
struct FFRows {	// form factor (both halves) of unique vertex u in cache row `row`
	const float2 *ff;
	int U;
	__host__ __device__ __forceinline__ float operator()(int row, int u) const {
		const float2 h = ldg(&ff[(size_t)row * U + u]);
		return h.x + h.y;
	}
};

// One CTA of 256 threads.  Candidate i = cache row i (the slot F.row_slot[i]); warp 0 runs the chain with lane l
// holding candidates l, l + 32, ... (the sequential restatement is chain_schedule_host in shot_core.cuh).
// Dynamic shared memory: rows x CandRec (the candidates' records with their rows x 3 table of form factors).
constexpr int kChainPerLane = kMaxRows / 32;
inline size_t chain_smem_bytes(int rows) { return (size_t)rows * sizeof(CandRec); }

// The chain for a single shooter per iteration, CPL = candidates per lane actually in use (rows <= 32 CPL): the loops over a
// lane's candidates are unrolled CPL times instead of kChainPerLane = 4 times (a third of the loop's instructions at
// rows <= 64).  State arrays stay in registers (forceinline, constant indices).
template <int CPL>
__device__ __forceinline__ void chain_single(const DevScene &S, const FastState &F, const CandRec *cm, int lane, float4 (&a0)[kChainPerLane],
					      float4 (&a1)[kChainPerLane], float4 (&a2)[kChainPerLane], const float (&fkx)[kChainPerLane],
					      const float (&fky)[kChainPerLane], const float (&fkz)[kChainPerLane], const int (&slot)[kChainPerLane],
					      unsigned long long (&key)[kChainPerLane]) {
	// single shooter per iteration: the chain stays in registers (one shuffle round per shot)
	const long long budget1 = *F.budget;
	const int kmax = (int)(budget1 < (long long)kMaxChain ? (budget1 < 0 ? 0 : budget1) : (long long)kMaxChain);
	auto lane_best = [&]() {
		unsigned long long m = 0ull;
#pragma unroll
		for (int c = 0; c < CPL; c++) m = key[c] > m ? key[c] : m;
		return m;
	};
	unsigned long long best = warp_max_u64(lane_best());
	int k = 0;
	while (best != 0ull && k < kmax) {
		// the shooter: the lane that holds the best key (keys carry the slot, so non-zero keys are unique) and which of its candidates
		int mine_c = -1;
#pragma unroll
		for (int c = 0; c < CPL; c++)
			if (key[c] == best) mine_c = c;
		const int bl = __ffs(__ballot_sync(0xffffffffu, mine_c >= 0)) - 1;
		float d0 = 0.f, d1 = 0.f, d2 = 0.f;
		int ss = -1;
#pragma unroll
		for (int c = 0; c < CPL; c++)
			if (c == mine_c) d0 = a0[c].w, d1 = a1[c].w, d2 = a2[c].w, ss = slot[c];
		const int bc = __shfl_sync(0xffffffffu, mine_c, bl);
		const int b = bl + 32 * bc;	// the shooter's candidate index = its cache row
		const int sslot = __shfl_sync(0xffffffffu, ss, bl);
		const float sd0 = __shfl_sync(0xffffffffu, d0, bl), sd1 = __shfl_sync(0xffffffffu, d1, bl), sd2 = __shfl_sync(0xffffffffu, d2, bl);
#pragma unroll
		for (int c = 0; c < CPL; c++) {
			if (slot[c] < 0) continue;
			if (slot[c] == sslot) a0[c].w = a1[c].w = a2[c].w = 0.0f;	// rad.cpp:446-449
			else if (unsaturated(a0[c], a1[c], a2[c])) {
				const float *f = cm[lane + 32 * c].fsum[b];
				apply_shot_ff(a0[c], a1[c], a2[c], fkx[c], fky[c], fkz[c], f[0], f[1], f[2], sd0, sd1, sd2);
			}
			key[c] = energy_key(S.converge, a0[c].w, a1[c].w, a2[c].w, slot[c]);
		}
		best = warp_max_u64(lane_best());
		if (lane == 0) {
			ShotSched s;
			s.slot = sslot, s.row = b, s.dB[0] = sd0, s.dB[1] = sd1, s.dB[2] = sd2, s.flags = kSchedHead | kSchedLast, s.next_key = best;
			F.sched[k] = s;
		}
		k++;
	}
	if (lane == 0) {
		*F.nsched = k, *F.viol = k, *F.apply_ticket = 0u;
		F.counters[2] += (unsigned long long)k;
		dbg_trace(F, 7);
	}
}

// 512 threads x 128 registers = the whole register file of an SM: no ff_batch_kernel CTA shares the SM with the chain,
// whose single warp otherwise got one issue slot in thirteen (measured 1.3 us per shot beside twelve ff_batch CTAs).
constexpr int kChainGrid = 32;		// CTAs that share the candidates' rays; the last one to finish runs the chain
constexpr int kChainBlock = 512;
__global__ void __launch_bounds__(kChainBlock, 1) chain_kernel(DevScene S, FastState F) {
	extern __shared__ __align__(16) unsigned char chain_smem[];
	CandRec *cm = (CandRec *)chain_smem;
	__shared__ int s_owned[kMaxRows];
	__shared__ int s_ok;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int rows = F.rows;
	// ---- phase 0 (all CTAs): the new rows' form factors at the candidates' vertices.  The kernel is launched on the side
	// branch of the round with launch priority, so its CTAs are resident BEFORE ff_batch_kernel fills every SM (a
	// stand-alone chain CTA - 128 registers per thread, ~80 KB of shared memory - did not fit beside twelve
	// ff_batch CTAs per SM and only started when ff_batch_kernel had handed out its last wave).
	{
		__shared__ bool s_last;
		if (blockIdx.x == 0 && threadIdx.x == 0) dbg_trace(F, 4);
		const int nt = *F.ntodo, per = rows * 6;
		for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < nt * per; w += gridDim.x * blockDim.x) ff_cand_item(S, F, w / per, w % per);
		__threadfence();
		__syncthreads();
		if (threadIdx.x == 0) s_last = atomicAdd(F.chain_ticket, 1u) == gridDim.x - 1;
		__syncthreads();
		if (!s_last) return;
		__threadfence();
		if (threadIdx.x == 0) *F.chain_ticket = 0u;
	}
	const unsigned long long round = *F.pred_round;
	const int par = (int)(round & 1ull);
	if (threadIdx.x == 0) s_ok = *F.error == 0, dbg_trace(F, 6);
	// own candidates: state, reflectance and the form factors of every filled row at their three vertices.  Three
	// dependent round trips in all: directory -> receiver -> (state, vertex ids) -> form factors of every (candidate, row).
	__shared__ int4 s_id[kMaxRows];
	__shared__ int s_filled[kMaxRows];
	for (int i = threadIdx.x; i < kMaxRows; i += blockDim.x) {
		const int slot = i < rows ? F.row_slot[i] : -1;
		const int r = slot >= 0 ? S.slot_to_recv[slot] : -1;
		s_owned[i] = r >= 0, s_filled[i] = slot >= 0;
		if (i < rows) cm[i].slot = r >= 0 ? slot : -1;	// free row, or a candidate owned by another rank (filled in below)
		if (r >= 0) {
			const int4 id = __ldg(&S.vid[r]);
			s_id[i] = id;
			cm[i].a0 = S.st[0][r], cm[i].a1 = S.st[1][r], cm[i].a2 = S.st[2][r];
			const float4 fk = __ldg(&S.par_fk[id.w]);
			cm[i].fk[0] = fk.x, cm[i].fk[1] = fk.y, cm[i].fk[2] = fk.z;
		}
	}
	__syncthreads();
	// rows filled in this round: ff_batch_kernel may still be writing them - their values at the candidates' vertices
	// come from ff_cand_kernel's table (s_todo_of[row] = todo entry, -1 for rows cached earlier)
	__shared__ int s_todo_of[kMaxRows];
	for (int i = threadIdx.x; i < kMaxRows; i += blockDim.x) s_todo_of[i] = -1;
	__syncthreads();
	{
		const int nt = *F.ntodo;
		for (int t = threadIdx.x; t < nt; t += blockDim.x) s_todo_of[F.todo[t]] = t;
	}
	__syncthreads();
	for (int q = threadIdx.x; q < rows * rows; q += blockDim.x) {	// (candidate i, row j)
		const int i = q / rows, j = q - i * rows;
		if (!s_owned[i]) continue;
		const bool filled = s_filled[j];
		const int4 id = s_id[i];
		const int t = s_todo_of[j];
		if (t >= 0) {
			const float2 *cf = F.cand_ff + ((size_t)t * kMaxRows + i) * 3;
			const float2 h0 = __ldcg(&cf[0]), h1 = __ldcg(&cf[1]), h2 = __ldcg(&cf[2]);
			cm[i].fsum[j][0] = h0.x + h0.y, cm[i].fsum[j][1] = h1.x + h1.y, cm[i].fsum[j][2] = h2.x + h2.y;
		} else {
			cm[i].fsum[j][0] = filled ? ff_sum(F.ff, S.U, j, id.x) : 0.0f;
			cm[i].fsum[j][1] = filled ? ff_sum(F.ff, S.U, j, id.y) : 0.0f;
			cm[i].fsum[j][2] = filled ? ff_sum(F.ff, S.U, j, id.z) : 0.0f;
		}
	}
	__syncthreads();
	// what travels of a record: the header and the first `rows` entries of its form-factor table
	const int rec_words = (int)(offsetof(CandRec, fsum) / 4) + 3 * rows;
	if (F.world > 1) {
		// publish the owned records to every other rank (4-byte granules, one warp per record), then the sequence number
		for (int i = warp; i < rows; i += (int)(blockDim.x >> 5)) {
			if (!s_owned[i]) continue;
			const int *src = (const int *)&cm[i];
			for (int t = 0; t < F.world; t++) {
				if (t == F.rank) continue;
				volatile int *dst = (volatile int *)&F.box[t]->cand[par][i];
				for (int w = lane; w < rec_words; w += 32) dst[w] = src[w];
			}
		}
		__syncthreads();
		if (threadIdx.x == 0) __threadfence_system();	// cumulative: the CTA's records before the sequence numbers
		__syncthreads();
		if (threadIdx.x < F.world && threadIdx.x != F.rank)
			*((volatile unsigned long long *)&F.box[threadIdx.x]->cand_seq[par][F.rank]) = round;
		PeerBox *mine = F.box[F.rank];
		if (threadIdx.x < F.world && threadIdx.x != F.rank && s_ok && !peer_wait(&mine->cand_seq[par][threadIdx.x], round)) s_ok = 0;
		__syncthreads();
		if (s_ok) {
			for (int i = warp; i < rows; i += (int)(blockDim.x >> 5)) {
				if (s_owned[i] || F.row_slot[i] < 0) continue;
				const volatile int *src = (const volatile int *)&mine->cand[par][i];
				int *dst = (int *)&cm[i];
				for (int w = lane; w < rec_words; w += 32) dst[w] = src[w];
			}
		}
		__syncthreads();
	}
	if (warp != 0) return;
	if (!s_ok) {
		if (lane == 0) *F.error = *F.error ? *F.error : 12, *F.nsched = 0, *F.viol = 0;	// 12: a rank's candidate records did not arrive
		return;
	}
	// ---- the chain: lane l replays the shots on candidates l + 32 c
	float4 a0[kChainPerLane], a1[kChainPerLane], a2[kChainPerLane];
	float fkx[kChainPerLane], fky[kChainPerLane], fkz[kChainPerLane];
	int slot[kChainPerLane];
	unsigned long long key[kChainPerLane];
#pragma unroll
	for (int c = 0; c < kChainPerLane; c++) {
		const int i = lane + 32 * c;
		const bool valid = i < rows && cm[i].slot >= 0;
		slot[c] = valid ? cm[i].slot : -1;
		if (valid) {
			a0[c] = cm[i].a0, a1[c] = cm[i].a1, a2[c] = cm[i].a2;
			fkx[c] = cm[i].fk[0], fky[c] = cm[i].fk[1], fkz[c] = cm[i].fk[2];
		} else {
			a0[c] = a1[c] = a2[c] = make_float4(0.f, 0.f, 0.f, 0.f);
			fkx[c] = fky[c] = fkz[c] = 0.f;
		}
		key[c] = valid ? energy_key(S.converge, a0[c].w, a1[c].w, a2[c].w, slot[c]) : 0ull;
	}
	if (F.parallel_src <= 1) {
		switch ((rows + 31) >> 5) {
		case 1: chain_single<1>(S, F, cm, lane, a0, a1, a2, fkx, fky, fkz, slot, key); break;
		case 2: chain_single<2>(S, F, cm, lane, a0, a1, a2, fkx, fky, fkz, slot, key); break;
		case 3: chain_single<3>(S, F, cm, lane, a0, a1, a2, fkx, fky, fkz, slot, key); break;
		default: chain_single<4>(S, F, cm, lane, a0, a1, a2, fkx, fky, fkz, slot, key); break;
		}
		return;
	}
	// One iteration of the reference's loop = n co-shooters (n = 1 without -parallel_src): the n largest keys, in
	// descending order (the sequential restatement is chain_schedule_host in shot_core.cuh).
	__shared__ ShotSched s_it[kMaxSources];
	const int P = F.parallel_src;
	long long lb = P > 1 ? *F.cand_count : 0ll;	// |C|: exact for the first iteration, a lower bound afterwards
	const long long budget = *F.budget;		// iterations this step may still run
	int k = 0, iters = 0, last = -1;
	for (;;) {
		if ((long long)iters >= budget) break;
		if (iters > 0 && !shooter_count_certain(lb, P)) break;
		const int n = P <= 1 ? 1 : shooters_for(lb, P);
		if (n <= 0 || k + n > kMaxChain) break;
		unsigned sel = 0u;	// bit c: my candidate c shoots in this iteration
		unsigned long long minkey = 0ull;
		bool ok = true;
		for (int j = 0; j < n; j++) {
			unsigned long long m = 0ull;
#pragma unroll
			for (int c = 0; c < kChainPerLane; c++)
				if (!((sel >> c) & 1u) && key[c] > m) m = key[c];
			const unsigned long long best = warp_max_u64(m);
			if (best == 0ull) { ok = false; break; }	// the candidates cannot supply n shooters (uniform)
			// the shooter: the lane that holds the best key (keys carry the slot, so non-zero keys are unique) and which of its candidates
			int mine_c = -1;
#pragma unroll
			for (int c = 0; c < kChainPerLane; c++)
				if (!((sel >> c) & 1u) && key[c] == best) mine_c = c;
			if (mine_c >= 0) {
				sel |= 1u << mine_c;
				ShotSched e;
				e.slot = -1, e.row = lane + 32 * mine_c;	// the shooter's candidate index = its cache row
				e.dB[0] = e.dB[1] = e.dB[2] = 0.f;
#pragma unroll
				for (int c = 0; c < kChainPerLane; c++)
					if (c == mine_c) e.slot = slot[c], e.dB[0] = a0[c].w, e.dB[1] = a1[c].w, e.dB[2] = a2[c].w;
				e.flags = (j == 0 ? kSchedHead : 0) | (j == n - 1 ? kSchedLast : 0);
				e.next_key = 0ull;
				s_it[j] = e;
			}
			minkey = best;
		}
		if (!ok) break;
		__syncwarp();
		if (last >= 0 && lane == 0) F.sched[last].next_key = minkey;
		for (int j = lane; j < n; j += 32) F.sched[k + j] = s_it[j];
#pragma unroll
		for (int c = 0; c < kChainPerLane; c++) {
			if (slot[c] < 0) continue;
			if ((sel >> c) & 1u) {	// rad.cpp:421-422: co-shooters do not receive; 446-449: their unshot energy is spent
				a0[c].w = a1[c].w = a2[c].w = 0.0f;
			} else if (unsaturated(a0[c], a1[c], a2[c])) {	// rad.cpp:424-428: once per iteration
				for (int j = 0; j < n; j++) {
					const float *f = cm[lane + 32 * c].fsum[s_it[j].row];
					apply_shot_ff(a0[c], a1[c], a2[c], fkx[c], fky[c], fkz[c], f[0], f[1], f[2], s_it[j].dB[0], s_it[j].dB[1], s_it[j].dB[2]);
				}
			}
			key[c] = energy_key(S.converge, a0[c].w, a1[c].w, a2[c].w, slot[c]);
		}
		__syncwarp();	// s_it is rewritten by the next iteration; the entries' stores are ordered before the next_key patch
		last = k + n - 1;
		k += n, iters++, lb -= n;
	}
	{
		// key of the shooter that would come next if the round went on (nothing scheduled follows: the test it feeds is harmless)
		unsigned long long m = 0ull;
#pragma unroll
		for (int c = 0; c < kChainPerLane; c++) m = key[c] > m ? key[c] : m;
		const unsigned long long best = warp_max_u64(m);
		if (lane == 0 && last >= 0) F.sched[last].next_key = P <= 1 ? best : ~0ull;
	}
	if (lane == 0) {
		*F.nsched = k, *F.viol = k, *F.apply_ticket = 0u;
		F.counters[2] += (unsigned long long)iters;
		dbg_trace(F, 7);
	}
}

// K3b for the whole schedule: one receiver per thread.  COMMIT = false: the apply pass (reads st, leaves the old state
// in st_prev, writes the state after all scheduled shots, reports the first violation).  COMMIT = true: after the
// first violation is known - re-applies the valid prefix from st_prev if the schedule went further, and does the
// round's bookkeeping.
// 128 threads x 7 CTAs per SM (72 registers): 896 receivers per SM in flight, so a church-sized shard (116 791 receivers,
// 789 per SM) is ONE wave; with 256 x 3 (80 registers, 768 per SM) a second, nearly empty wave doubled the pass
// (measured church update 0.87 -> 0.77 ms; 256 x 4: 0.82, 128 x 6: 0.84; box / room unchanged)
#ifndef RADK_APPLY_BLOCK
#define RADK_APPLY_BLOCK 128
#endif
#ifndef RADK_APPLY_MINB
#define RADK_APPLY_MINB 7
#endif
constexpr int kApplyBlock = RADK_APPLY_BLOCK;
template <bool COMMIT, bool MULTI>	// MULTI: -parallel_src > 1 (iterations with co-shooters)
__global__ void __launch_bounds__(kApplyBlock, RADK_APPLY_MINB) apply_kernel(DevScene S, FastState F) {
	__shared__ __align__(32) ShotSched sched[kMaxChain];
	__shared__ int s_n, s_wmin[kApplyBlock / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int K = *F.nsched;
	int n = K;
	if (blockIdx.x == 0 && threadIdx.x == 0) dbg_trace(F, COMMIT ? 9 : 8);
	if (COMMIT) {
		if (threadIdx.x == 0) {
			int v = *((volatile int *)F.viol);
			if (F.world > 1 && K > 0) {	// K is the same on every rank (identical chains): nobody mails an empty round
				// first violation over all ranks: every CTA reads the ranks' mail from its own box
				const unsigned long long round = *F.pred_round;
				const int par = (int)(round & 1ull);
				const PeerBox *mine = F.box[F.rank];
				const unsigned long long want = viol_word(round, 0) >> 8;
				for (int t = 0; t < F.world; t++) {
					if (t == F.rank) continue;
					const volatile unsigned long long *vp = &mine->viol[par][t];
					const long long t0 = clock64();
					unsigned spins = 0;
					unsigned long long w;
					while (((w = *vp) >> 8) != want) {
						if ((++spins & 0xfff) == 0 && clock64() - t0 > 40000000000ll) { *F.error = 13; w = 0; break; }	// 13: a rank's first violation did not arrive
					}
					const int vt = (int)(w & 0xffull);
					v = vt < v ? vt : v;
				}
				if (*((volatile int *)F.error)) v = -1;	// a peer is gone: commit nothing
			}
			s_n = v < K ? v + 1 : K;
		}
		__syncthreads();
		n = s_n;
		if (blockIdx.x == 0) {	// bookkeeping of the round (rad.cpp:504-514: loop counter; shot log for the tests)
			// an iteration counts once and logs its first shooter (n ends an iteration: violations are reported on last entries)
			__shared__ int s_iters;
			const long long before = *S.sel.shots_done;
			if (!MULTI) {
				for (int k = threadIdx.x; k < n; k += kApplyBlock) S.sel.shot_log[(before + k) % S.sel.shot_log_cap] = F.sched[k].slot;
				if (threadIdx.x == 0) s_iters = n;
			} else if (threadIdx.x == 0) {
				int it = 0;
				for (int k = 0; k < n; k++)
					if (F.sched[k].flags & kSchedHead) S.sel.shot_log[(before + it++) % S.sel.shot_log_cap] = F.sched[k].slot;
				s_iters = it;
			}
			__syncthreads();
			if (threadIdx.x == 0) {
				const int it = s_iters;
				*S.sel.shots_done = before + it;
				*F.budget -= it;
				atomicAdd(&S.stats->shots, (unsigned long long)it);
				F.counters[3] += (unsigned long long)it;
				if (n < K) F.counters[4]++;
				// device-side round loop (CUDA graph while node): another round unless converged, out of budget or failed
				if (F.cond) {
					const bool more = n > 0 && !*S.sel.converged && *F.budget > 0 && !*((volatile int *)F.error);
					cudaGraphSetConditional((cudaGraphConditionalHandle)F.cond, more ? 1u : 0u);
				}
				if (F.dbg) dbg_trace(F, 10), F.dbg[15]++;
			}
		}
		if (n >= K) return;	// the whole schedule was valid: the apply pass already wrote the right state
	}
	if (n <= 0) return;
#if RADK_TMA_APPLY
	{	// the round's schedule (n x 32 bytes) by one bulk copy
		__shared__ __align__(8) unsigned long long s_bar;
		if (threadIdx.x == 0) {
			mbar_init(&s_bar, 1);
			mbar_expect_tx(&s_bar, (unsigned)(n * sizeof(ShotSched)));
			tma_load_1d(sched, F.sched, (unsigned)(n * sizeof(ShotSched)), &s_bar);
		}
		__syncthreads();
		mbar_wait(&s_bar, 0u);
	}
#else
	for (int k = threadIdx.x; k < n; k += kApplyBlock) sched[k] = F.sched[k];
	__syncthreads();
#endif
	const int r = blockIdx.x * kApplyBlock + threadIdx.x;
	int viol = n;
	if (r < S.R) {
		float4 a0, a1, a2;
		if (COMMIT) {
			a0 = F.st_prev[0][r], a1 = F.st_prev[1][r], a2 = F.st_prev[2][r];
		} else {
			a0 = S.st[0][r], a1 = S.st[1][r], a2 = S.st[2][r];
			F.st_prev[0][r] = a0, F.st_prev[1][r] = a1, F.st_prev[2][r] = a2;
		}
		const int4 id = __ldg(&S.vid[r]);
		const int slot = __ldg(&S.slot[r]);
		const float4 fk = __ldg(&S.par_fk[id.w]);
		if (MULTI) viol = apply_schedule_multi(a0, a1, a2, id.x, id.y, id.z, fk.x, fk.y, fk.z, slot, sched, n, S.converge, FFRows{F.ff, S.U});
		else viol = apply_schedule_single(a0, a1, a2, id.x, id.y, id.z, fk.x, fk.y, fk.z, slot, sched, n, S.converge, FFRows{F.ff, S.U});
		S.st[0][r] = a0, S.st[1][r] = a1, S.st[2][r] = a2;
	}
	if (COMMIT) return;
	// first violation of this rank: CTA minimum -> atomicMin; multi-GPU: the last CTA mails the rank's value
	const int wmin = __reduce_min_sync(0xffffffffu, viol);
	if (lane == 0) s_wmin[warp] = wmin;
	__syncthreads();
	if (threadIdx.x == 0) {
		int m = n;
		for (int w = 0; w < kApplyBlock / 32; w++) m = s_wmin[w] < m ? s_wmin[w] : m;
		if (m < n) atomicMin(F.viol, m);
		if (F.world > 1) {
			__threadfence();
			if (atomicAdd(F.apply_ticket, 1u) == gridDim.x - 1) {
				__threadfence();
				const int v = *((volatile int *)F.viol);
				const unsigned long long round = *F.pred_round;
				const unsigned long long w = viol_word(round, v);
				for (int t = 0; t < F.world; t++)
					if (t != F.rank) *((volatile unsigned long long *)&F.box[t]->viol[(int)(round & 1ull)][F.rank]) = w;
			}
		}
	}
}

}  // namespace radk
