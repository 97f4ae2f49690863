// Device context of the shooting loop: scene upload (flattening the octree, sorting receivers,
// sharing vertices), the shot loop, downloads, and the C ABI declared in include/radk.h.
//
// Everything that computes runs in the kernels of shoot_kernels.cuh; there is no CPU fallback -
// every entry point fails with RADK_ERR_CUDA when no device is usable.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <numeric>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "radk.h"
#include "shoot_kernels.cuh"
#include "fast_kernels.cuh"
#include "split_kernels.cuh"
#include "scene_flatten.hpp"

namespace radk {
void set_error(const std::string &msg);	// host/model_capi.cpp
}

using namespace radk;

#define CUDA_TRY(expr)                                                                             \
	do {                                                                                       \
		cudaError_t e_ = (expr);                                                           \
		if (e_ != cudaSuccess) {                                                           \
			set_error(std::string(#expr) + ": " + cudaGetErrorString(e_));             \
			return RADK_ERR_CUDA;                                                      \
		}                                                                                  \
	} while (0)

namespace {

template <typename T> struct DevBuf {
	T *p = nullptr;
	size_t n = 0;
	cudaError_t alloc(size_t count) {
		release();
		n = count;
		return cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
	}
	cudaError_t upload(const std::vector<T> &h, cudaStream_t s) {
		cudaError_t e = alloc(h.size());
		if (e != cudaSuccess || h.empty()) return e;
		return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
	}
	void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
	~DevBuf() { release(); }
};

}  // namespace

struct radk_ctx {
	radk_opts opt{};
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	cudaEvent_t stage_ev[4] = {nullptr, nullptr, nullptr, nullptr};	// kStageChunks
	double device_ms = 0;
	long long launches = 0, launches_per_round = 5;

	int trinum = 0, n0 = 0;
	int R = 0, U = 0, leaves_total = 0;
	std::vector<int> recv_slot;		// receiver index -> slot (this shard)
	std::vector<int> recv_parent;
	std::vector<int> h_slot_to_recv;	// slot -> receiver index on this shard, -1 otherwise

	DevBuf<int> cell_leaf, list, slot_to_recv, shot_log, slot, leaf_list4, tri_leaf_begin, tri_leaves, uv_leaf;
	DevBuf<int4> list4;
	DevBuf<unsigned> row_mask;
	DevBuf<float4> leaf_box, occ, geo[3], st[3], st0[3], uv, par_n, par_fk;
	DevBuf<int4> vid;
	DevBuf<float2> ffh;
	DevBuf<unsigned> vstat, partial_cnt, ticket;
	DevBuf<unsigned long long> partial_key, bound;
	DevBuf<int> flags;			// [0]=ncand [1]=nsrc [2]=converged
	DevBuf<long long> shots_done;
	DevBuf<DevStats> stats;
	DevBuf<ShooterRec> rec;			// [0],[1]: ping-pong for the single-source loop; [2..]: multi-source list
	// fast path (fast_kernels.cuh)
	DevBuf<float2> ffcache;
	DevBuf<int> row_slot, slot_row, todo, fast_ints;	// fast_ints: [0]=ntodo [1]=nsched [2]=viol
	DevBuf<ShooterRec> row_rec;
	DevBuf<unsigned long long> group_key;
	DevBuf<unsigned> fast_uints;			// [0]=ticket [1]=apply ticket
	DevBuf<ShotRec> handoff;
	DevBuf<ShotSched> sched;
	DevBuf<float4> st_prev[3];
	DevBuf<unsigned long long> counters;
	DevBuf<long long> budget;
	DevBuf<PeerBox> box;
	DevBuf<unsigned long long> pred_round;
	DevBuf<int> fast_error;
	DevBuf<unsigned long long> exact;	// -parallel_src: threshold, list length, collected keys, sorted group maxima
	DevBuf<long long> cand_count;
	DevBuf<float2> cand_ff;
	cudaStream_t side = nullptr;		// ff_cand_kernel + chain_kernel run here, beside ff_batch_kernel (host-launched rounds)
	bool overlap = true;			// false: ranks share this device (tests) - one stream per rank, see radk_ctx_peer_connect_local
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	DevBuf<unsigned long long> dbg;
	std::vector<void *> ipc_opened;
	bool peers_connected = false;
	unsigned long long epoch = 0, epoch_word = 0;
	cudaGraph_t loop_graph = nullptr;	// while (more) { one round }: the whole step is one launch
	cudaGraphExec_t loop_exec = nullptr;
	bool graph_tried = false;
	int *h_error = nullptr;			// pinned
	FastState F{};
	bool use_fast = false;
	int rows = 0;
	long long *h_budget = nullptr;		// pinned
	int max_blocks = 0;
	int cur = 0;				// which of rec[0]/rec[1] holds the current shooter
	bool have_current = false;		// rec[cur] is up to date with the device state
	DevScene S{};
	struct HostFlags { int ncand, nsrc, converged; } *h_flags = nullptr;	// pinned
	long long *h_shots = nullptr;		// pinned
	float4 *h_stage = nullptr;		// pinned staging for the state planes: 3 x R float4
	bool use_dedup = true;
	// adaptive-split mode (split_kernels.cuh): slot-indexed pool instead of the sorted receiver planes
	bool split_mode = false;
	SplitScene SP{};
	DevBuf<float4> sp_p[3], sp_n, sp_st[3], sp_fk, sp_ffs;
	DevBuf<radk_triangle> sp_rec;
	DevBuf<int> sp_count, sp_bsum, sp_boff, sp_trinum, sp_growth;
	int sp_host_trinum = 0;
	unsigned long long sp_shot_no = 0;
};

namespace {

const int kShotLogCap = 1 << 20;
const int kBatchShots = 32;

// Host-side gather / scatter between the caller's TriStore-shaped buffers and the pinned staging planes is a
// strided walk over hundreds of MB on large scenes: a small persistent pool of host threads shares it
// (RADK_HOST_THREADS, default min(8, cores); small jobs stay on the calling thread).
class HostPool {
public:
	static HostPool &get() { static HostPool p; return p; }
	template <typename F> void run(size_t n, F fn) {
		const int nt = (int)std::min<size_t>((size_t)workers_.size() + 1, n / 8192 + 1);
		if (nt <= 1) { fn((size_t)0, n); return; }
		std::lock_guard<std::mutex> one_caller(run_m_);	// contexts driven from different host threads take turns
		const size_t chunk = (n + nt - 1) / nt;
		std::function<void(int)> job = [&](int k) { fn(std::min(n, k * chunk), std::min(n, (k + 1) * chunk)); };
		{
			std::lock_guard<std::mutex> g(m_);
			job_ = &job, parts_ = nt, next_ = 1, pending_ = nt - 1, gen_++;
		}
		cv_.notify_all();
		job(0);
		std::unique_lock<std::mutex> g(m_);
		done_.wait(g, [&] { return pending_ == 0; });
		job_ = nullptr;
	}
private:
	HostPool() {
		const char *e = getenv("RADK_HOST_THREADS");
		int t = e ? atoi(e) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
		t = std::max(1, std::min(t, 64));
		for (int k = 1; k < t; k++) workers_.emplace_back([this] { loop(); });
	}
	~HostPool() {
		{ std::lock_guard<std::mutex> g(m_); stop_ = true; }
		cv_.notify_all();
		for (auto &t : workers_) t.join();
	}
	void loop() {
		unsigned long long seen = 0;
		std::unique_lock<std::mutex> g(m_);
		for (;;) {
			cv_.wait(g, [&] { return stop_ || (gen_ != seen && job_ && next_ < parts_); });
			if (stop_) return;
			while (job_ && next_ < parts_) {
				const int k = next_++;
				std::function<void(int)> *job = job_;
				g.unlock();
				(*job)(k);
				g.lock();
				if (--pending_ == 0) done_.notify_all();
			}
			seen = gen_;
		}
	}
	std::vector<std::thread> workers_;
	std::mutex m_, run_m_;
	std::condition_variable cv_, done_;
	std::function<void(int)> *job_ = nullptr;
	int parts_ = 0, next_ = 0, pending_ = 0;
	unsigned long long gen_ = 0;
	bool stop_ = false;
};
template <typename F> void host_parallel_for(size_t n, F fn) { HostPool::get().run(n, fn); }
constexpr int kStageChunks = 4;	// upload / download are pipelined in this many receiver ranges

int reset_fast_state(radk_ctx *c) {
	if (!c->use_fast) return RADK_OK;
	cudaStream_t st = c->stream;
	CUDA_TRY(cudaMemsetAsync(c->row_slot.p, 0xff, kMaxRows * sizeof(int), st));
	CUDA_TRY(cudaMemsetAsync(c->slot_row.p, 0xff, std::max(c->trinum, 1) * sizeof(int), st));
	CUDA_TRY(cudaMemsetAsync(c->todo.p, 0, kMaxRows * sizeof(int), st));
	CUDA_TRY(cudaMemsetAsync(c->fast_ints.p, 0, 4 * sizeof(int), st));
	CUDA_TRY(cudaMemsetAsync(c->group_key.p, 0, kPredictGroups * sizeof(unsigned long long), st));
	CUDA_TRY(cudaMemsetAsync(c->fast_uints.p, 0, 4 * sizeof(unsigned), st));
	CUDA_TRY(cudaMemsetAsync(c->handoff.p, 0, sizeof(ShotRec), st));
	CUDA_TRY(cudaMemsetAsync(c->budget.p, 0, sizeof(long long), st));
	CUDA_TRY(cudaMemsetAsync(c->fast_error.p, 0, sizeof(int), st));
	CUDA_TRY(cudaMemsetAsync(c->counters.p, 0, 8 * sizeof(unsigned long long), st));
	CUDA_TRY(cudaMemsetAsync(c->exact.p, 0, (2 + kExactCap + kMaxRows) * sizeof(unsigned long long), st));
	CUDA_TRY(cudaMemsetAsync(c->cand_count.p, 0, sizeof(long long), st));
	c->epoch_word = (++c->epoch) << 40;	// stale mailbox sequence numbers of earlier runs can never match
	CUDA_TRY(cudaMemcpyAsync(c->pred_round.p, &c->epoch_word, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
	return RADK_OK;
}

// CUDA loads a kernel's code at its first launch (lazy module loading), and that load can wait for kernels already
// running on the device.  Kernels of this path WAIT for their peers' mail: with several ranks on one device (tests,
// `rad -gpus N` on fewer devices) a rank's first launch of, say, predict_exact_kernel would sit behind the other rank's
// spinning merge kernel until the device-side timeout.  So every kernel of the round is loaded when a context is created.
int preload_round_kernels() {
	const void *fns[] = {(const void *)predict_kernel, (const void *)predict_exact_kernel, (const void *)predict_merge_kernel,
			     (const void *)ff_batch_kernel<1>, (const void *)ff_batch_kernel<2>, (const void *)ff_batch_kernel<4>,
			     (const void *)chain_kernel, (const void *)apply_kernel<false, false>, (const void *)apply_kernel<true, false>,
			     (const void *)apply_kernel<false, true>, (const void *)apply_kernel<true, true>};
	for (const void *f : fns) {
		cudaFuncAttributes a;
		CUDA_TRY(cudaFuncGetAttributes(&a, f));
	}
	return RADK_OK;
}

// One round of the fast path: predict -> batched form factors -> chain -> apply -> commit.
// ev (optional): 4 events recorded around the three kernel classes, for profiling.
int launch_fast_round(radk_ctx *c, cudaEvent_t *ev) {
	cudaStream_t st = c->stream;
	FastState F = c->F;
	F.cond = 0;	// host-launched rounds: no graph condition to set
	if (ev) CUDA_TRY(cudaEventRecord(ev[0], st));
	predict_kernel<<<64, 1024, 0, st>>>(c->S, F);
	if (c->F.parallel_src > 1) { predict_exact_kernel<<<64, 1024, 0, st>>>(c->S, F); c->launches++; }
	if (c->F.world > 1) { predict_merge_kernel<<<1, 1024, 0, st>>>(c->S, F); c->launches++; }
	if (ev) CUDA_TRY(cudaEventRecord(ev[1], st));
	// fork: the form factors of all vertices on the main stream; beside them the candidates' own form factors and
	// the chain (a few thousand rays, then a single warp)
	// the side stream exists only where it is used: every stream takes one of the device's few hardware queues, and
	// ranks that share a device (tests) deadlock when a kernel waiting for a peer's mail is queued in front of the
	// peer's kernel that sends it
	if (c->overlap && !c->side) {
		int lo = 0, hi = 0;
		CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));	// greatest priority: the side branch's few CTAs go first
		CUDA_TRY(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi));
	}
	cudaStream_t sd = c->overlap ? c->side : st;
	if (c->overlap) {
		CUDA_TRY(cudaEventRecord(c->ev_fork, st));
		CUDA_TRY(cudaStreamWaitEvent(sd, c->ev_fork, 0));
	}
	chain_kernel<<<kChainGrid, kChainBlock, chain_smem_bytes(c->rows), sd>>>(c->S, F);
	if (c->overlap) CUDA_TRY(cudaEventRecord(c->ev_join, sd));
	const int items = ff_items_for(c->U, c->rows);
	const dim3 g((unsigned)std::max(1, (c->U + kFFBlock * items - 1) / (kFFBlock * items)), 2, (unsigned)c->rows);
	{
		void *args[] = {(void *)&c->S, (void *)&F};
		CUDA_TRY(cudaLaunchKernel(ff_batch_fn(items), g, dim3(kFFBlock), args, 0, st));
	}
	if (c->overlap) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_join, 0));
	if (ev) CUDA_TRY(cudaEventRecord(ev[2], st));
	const int nb = std::max(1, (c->R + kApplyBlock - 1) / kApplyBlock);
	if (c->F.parallel_src > 1) {
		apply_kernel<false, true><<<nb, kApplyBlock, 0, st>>>(c->S, F);
		apply_kernel<true, true><<<nb, kApplyBlock, 0, st>>>(c->S, F);
	} else {
		apply_kernel<false, false><<<nb, kApplyBlock, 0, st>>>(c->S, F);
		apply_kernel<true, false><<<nb, kApplyBlock, 0, st>>>(c->S, F);
	}
	if (ev) CUDA_TRY(cudaEventRecord(ev[3], st));
	c->launches += 5;
	return RADK_OK;
}

// The round loop as a CUDA graph with a device-side while node: predict -> (merge) -> form factors -> chain -> apply ->
// commit, repeated until the commit kernel clears the condition (converged / budget spent / error).  Built once per
// context (the kernel arguments never change: everything that does lives in device memory).  Returns false when the
// graph cannot be built; the caller then launches the rounds from the host.
bool build_loop_graph(radk_ctx *c) {
	if (c->graph_tried) return c->loop_exec != nullptr;
	c->graph_tried = true;
	if (getenv("RADK_NO_GRAPH")) return false;
	cudaGraph_t g = nullptr;
	if (cudaGraphCreate(&g, 0) != cudaSuccess) { cudaGetLastError(); return false; }
	cudaGraphConditionalHandle h;
	if (cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault) != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(g); return false; }
	cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
	cp.conditional.handle = h, cp.conditional.type = cudaGraphCondTypeWhile, cp.conditional.size = 1;
	cudaGraphNode_t cnode;
	if (cudaGraphAddNode(&cnode, g, nullptr, 0, &cp) != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(g); return false; }
	cudaGraph_t body = cp.conditional.phGraph_out[0];
	c->F.cond = (unsigned long long)h;
	void *args[] = {(void *)&c->S, (void *)&c->F};
	struct K { const void *fn; dim3 grid, block; size_t smem; int dep0, dep1; };	// dep: index of a kernel this one waits for (-1: none)
	const int nb = std::max(1, (c->R + kApplyBlock - 1) / kApplyBlock);
	std::vector<K> ks;
	auto add = [&](const void *fn, dim3 grid, dim3 block, size_t smem, int dep0, int dep1 = -1) {
		ks.push_back({fn, grid, block, smem, dep0, dep1});
		return (int)ks.size() - 1;
	};
	int last = add((const void *)predict_kernel, dim3(64), dim3(1024), 0, -1);
	if (c->F.parallel_src > 1) last = add((const void *)predict_exact_kernel, dim3(64), dim3(1024), 0, last);
	if (c->F.world > 1) last = add((const void *)predict_merge_kernel, dim3(1), dim3(1024), 0, last);
	// fork: the form factors of all vertices | the candidates' own form factors + the chain (chain_kernel, added first:
	// its few CTAs must be resident before ff_batch_kernel fills the device)
	const int kchain = add((const void *)chain_kernel, dim3(kChainGrid), dim3(kChainBlock), chain_smem_bytes(c->rows), last);
	const int items = ff_items_for(c->U, c->rows);
	const int kff = add(ff_batch_fn(items), dim3((unsigned)std::max(1, (c->U + kFFBlock * items - 1) / (kFFBlock * items)), 2, (unsigned)c->rows), dim3(kFFBlock), 0, last);
	if (c->F.parallel_src > 1) {
		last = add((const void *)apply_kernel<false, true>, dim3(nb), dim3(kApplyBlock), 0, kff, kchain);
		last = add((const void *)apply_kernel<true, true>, dim3(nb), dim3(kApplyBlock), 0, last);
	} else {
		last = add((const void *)apply_kernel<false, false>, dim3(nb), dim3(kApplyBlock), 0, kff, kchain);
		last = add((const void *)apply_kernel<true, false>, dim3(nb), dim3(kApplyBlock), 0, last);
	}
	std::vector<cudaGraphNode_t> nodes(ks.size());
	bool ok = true;
	for (size_t i = 0; i < ks.size(); i++) {
		const K &k = ks[i];
		cudaKernelNodeParams kp;
		memset(&kp, 0, sizeof(kp));
		kp.func = (void *)k.fn, kp.gridDim = k.grid, kp.blockDim = k.block, kp.sharedMemBytes = (unsigned)k.smem, kp.kernelParams = args, kp.extra = nullptr;
		cudaGraphNode_t deps[2];
		int nd = 0;
		if (k.dep0 >= 0) deps[nd++] = nodes[k.dep0];
		if (k.dep1 >= 0) deps[nd++] = nodes[k.dep1];
		if (cudaGraphAddKernelNode(&nodes[i], body, nd ? deps : nullptr, nd, &kp) != cudaSuccess) { ok = false; break; }
		if ((int)i == kchain) {
			// the side branch goes first: its CTAs are taken ahead of the thousands ff_batch_kernel still has pending
			// (without the priority it only ran when ff_batch_kernel's last wave had been handed out)
			int lo = 0, hi = 0;
			cudaDeviceGetStreamPriorityRange(&lo, &hi);	// hi = numerically lowest = greatest priority
			cudaKernelNodeAttrValue v;
			memset(&v, 0, sizeof(v));
			v.priority = hi;
			if (cudaGraphKernelNodeSetAttribute(nodes[i], cudaKernelNodeAttributePriority, &v) != cudaSuccess) cudaGetLastError();
		}
	}
	cudaGraphExec_t exec = nullptr;
	if (ok && cudaGraphInstantiate(&exec, g, 0) != cudaSuccess) ok = false;
	if (!ok) {
		cudaGetLastError();
		cudaGraphDestroy(g);
		c->F.cond = 0;
		return false;
	}
	c->loop_graph = g, c->loop_exec = exec;
	c->launches_per_round = (long long)ks.size();
	return true;
}

// ---- adaptive-split mode -----------------------------------------------------------------------------
int upload_split_scene(radk_ctx *c, const radk_scene *s) {
	const radk_opts &o = c->opt;
	cudaStream_t st = c->stream;
	c->split_mode = true;
	c->use_dedup = false, c->use_fast = false;
	c->n0 = s->n0;
	const int max_tri = o.max_tri > 0 ? o.max_tri : 1000000;
	const int limit = std::min(max_tri, o.tri_limit);
	if (s->trinum > limit) { set_error("scene larger than the triangle limit"); return RADK_ERR_ARG; }
	const int cap = limit;
	FlatScene F;
	radk_scene empty = *s;
	empty.trinum = 0;		// only the octree / occluder part of the flattening is needed
	const int frc = flatten_scene(&empty, 0, 1, F);
	if (frc != RADK_OK) { set_error(F.error); return frc; }
	CUDA_TRY(c->cell_leaf.upload(F.cell_leaf, st));
	CUDA_TRY(c->leaf_box.upload(F.leaf_box, st));
	CUDA_TRY(c->list.upload(F.list, st));
	CUDA_TRY(c->list4.upload(F.list4, st));
	CUDA_TRY(c->leaf_list4.upload(F.leaf_list4, st));
	CUDA_TRY(c->tri_leaf_begin.upload(F.tri_leaf_begin, st));
	CUDA_TRY(c->tri_leaves.upload(F.tri_leaves, st));
	CUDA_TRY(c->occ.upload(F.occ, st));
	std::vector<float4> p[3], stt[3], nn(cap), fk(cap);
	for (int v = 0; v < 3; v++) p[v].assign(cap, make_float4(0, 0, 0, 0)), stt[v].assign(cap, make_float4(0, 0, 0, 0));
	std::vector<radk_triangle> rec(cap);
	memset(rec.data(), 0, sizeof(radk_triangle) * (size_t)cap);
	const float kref = (float)(0.5 / 255.0);
	for (int i = 0; i < cap; i++) p[0][i].w = flat_detail::ibits(-1);
	for (int i = 0; i < s->trinum; i++) {
		const radk_triangle &t = s->tri[i];
		rec[i] = t;
		p[0][i] = make_float4(t.p[0][0], t.p[0][1], t.p[0][2], flat_detail::ibits(t.parent));
		p[1][i] = make_float4(t.p[1][0], t.p[1][1], t.p[1][2], t.area);
		p[2][i] = make_float4(t.p[2][0], t.p[2][1], t.p[2][2], 0.f);
		nn[i] = make_float4(t.n[0], t.n[1], t.n[2], sqrtf(t.n[0] * t.n[0] + t.n[1] * t.n[1] + t.n[2] * t.n[2]));
		fk[i] = make_float4(t.Frgb[0] * kref, t.Frgb[1] * kref, t.Frgb[2] * kref, 0.f);
		for (int v = 0; v < 3; v++) stt[v][i] = make_float4(t.accB[v][0], t.accB[v][1], t.accB[v][2], t.deltaB[v]);
	}
	for (int v = 0; v < 3; v++) {
		CUDA_TRY(c->sp_p[v].upload(p[v], st));
		CUDA_TRY(c->sp_st[v].upload(stt[v], st));
	}
	CUDA_TRY(c->sp_n.upload(nn, st));
	CUDA_TRY(c->sp_fk.upload(fk, st));
	CUDA_TRY(c->sp_rec.upload(rec, st));
	CUDA_TRY(c->sp_ffs.alloc(cap));
	CUDA_TRY(c->sp_count.alloc(cap));
	const int maxblocks = (cap + kScanBlock - 1) / kScanBlock;
	CUDA_TRY(c->sp_bsum.alloc(maxblocks));
	CUDA_TRY(c->sp_boff.alloc(maxblocks));
	CUDA_TRY(c->sp_trinum.alloc(1));
	CUDA_TRY(c->sp_growth.alloc(1));
	CUDA_TRY(cudaMemsetAsync(c->sp_count.p, 0, (size_t)cap * sizeof(int), st));
	c->sp_host_trinum = s->trinum;
	c->trinum = s->trinum;
	CUDA_TRY(cudaMemcpyAsync(c->sp_trinum.p, &c->sp_host_trinum, sizeof(int), cudaMemcpyHostToDevice, st));
	c->max_blocks = std::max(1, std::min((cap + 255) / 256, 4 * 148));
	CUDA_TRY(c->partial_key.alloc(c->max_blocks));
	CUDA_TRY(c->partial_cnt.alloc(c->max_blocks));
	CUDA_TRY(c->ticket.alloc(1));
	CUDA_TRY(c->bound.alloc(1));
	CUDA_TRY(c->flags.alloc(4));
	CUDA_TRY(c->shots_done.alloc(1));
	CUDA_TRY(c->shot_log.alloc(kShotLogCap));
	CUDA_TRY(c->stats.alloc(1));
	CUDA_TRY(c->rec.alloc(2 + kMaxSources));
	CUDA_TRY(cudaMemsetAsync(c->ticket.p, 0, sizeof(unsigned), st));
	CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, 4 * sizeof(int), st));
	CUDA_TRY(cudaMemsetAsync(c->shots_done.p, 0, sizeof(long long), st));
	CUDA_TRY(cudaMemsetAsync(c->stats.p, 0, sizeof(DevStats), st));
	CUDA_TRY(cudaMemsetAsync(c->rec.p, 0, (2 + kMaxSources) * sizeof(ShooterRec), st));
	SplitScene &S = c->SP;
	S.tree.cell_leaf = c->cell_leaf.p, S.tree.leaf_box = c->leaf_box.p;
	S.tree.list = c->list.p, S.tree.list4 = c->list4.p, S.tree.leaf_list4 = c->leaf_list4.p;
	S.tree.tri_leaf_begin = c->tri_leaf_begin.p, S.tree.tri_leaves = c->tri_leaves.p, S.tree.occ = c->occ.p, S.tree.nocc = s->n0;
	S.tree.nnodes = s->nnodes;
	for (int k = 0; k < 3; k++) S.tree.G0[k] = s->G0[k], S.tree.G1[k] = s->G1[k];
	S.tree.min_grid_len = s->min_grid_len, S.tree.grid_num = s->grid_num, S.tree.inv_min_grid_len = 1.0 / (double)s->min_grid_len;
	S.tree.inv_min_grid_len_f = (float)S.tree.inv_min_grid_len, S.tree.cell_eps = 4e-7f * (float)(s->grid_num + 1) + 1e-4f;
	S.capacity = cap, S.limit = limit, S.trinum = c->sp_trinum.p;
	for (int v = 0; v < 3; v++) S.p[v] = c->sp_p[v].p, S.st[v] = c->sp_st[v].p;
	S.n = c->sp_n.p, S.fk = c->sp_fk.p, S.rec = c->sp_rec.p, S.count = c->sp_count.p, S.ffs = c->sp_ffs.p;
	S.block_sum = c->sp_bsum.p, S.block_off = c->sp_boff.p;
	S.converge = o.converge, S.area_limit = o.area_limit, S.sample_area = o.sample_area, S.delta_ff = o.delta_ff;
	S.stats = c->stats.p;
	S.sel.partial_key = c->partial_key.p, S.sel.partial_cnt = c->partial_cnt.p, S.sel.ticket = c->ticket.p;
	S.sel.bound = c->bound.p;
	S.sel.ncand = c->flags.p, S.sel.nsrc = c->flags.p + 1, S.sel.converged = c->flags.p + 2;
	S.sel.shots_done = c->shots_done.p, S.sel.shot_log = c->shot_log.p, S.sel.shot_log_cap = kShotLogCap;
	S.shot_no = 0;
	if (!(o.sample_area > 0)) { set_error("adaptive splitting needs opts.sample_area > 0 (radk_model_sample_area)"); return RADK_ERR_ARG; }
	CUDA_TRY(cudaStreamSynchronize(st));
	return RADK_OK;
}

// one shot of the split-enabled loop (rad.cpp:406-451 with the do/while over the growing pool)
int split_one_shot(radk_ctx *c, bool *converged) {
	cudaStream_t st = c->stream;
	ShooterRec *sh = c->rec.p;
	split_select_kernel<<<c->max_blocks, 256, 0, st>>>(c->SP, sh);
	c->launches++;
	CUDA_TRY(cudaMemcpyAsync(c->h_flags, c->flags.p, sizeof(*c->h_flags), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	if (c->h_flags->converged) { *converged = true; return RADK_OK; }
	c->SP.shot_no = ++c->sp_shot_no;
	int prev = 0, n = c->sp_host_trinum;
	while (n > prev) {
		const int count = n - prev;
		const int nb = (count + kScanBlock - 1) / kScanBlock;
		split_shade_kernel<<<(count + kSplitBlock / kSplitGroup - 1) / (kSplitBlock / kSplitGroup), kSplitBlock, 0, st>>>(c->SP, sh, prev, n);
		split_scan_sums_kernel<<<nb, kScanBlock, 0, st>>>(c->SP, prev, n);
		split_scan_offsets_kernel<<<1, kScanBlock, 0, st>>>(c->SP, nb, c->sp_growth.p);
		split_append_kernel<<<nb, kScanBlock, 0, st>>>(c->SP, sh, prev, n, n);
		split_reshade_kernel<<<(count * 4 + 127) / 128, 128, 0, st>>>(c->SP, sh, prev, n);
		c->launches += 5;
		int growth = 0;
		CUDA_TRY(cudaMemcpyAsync(&growth, c->sp_growth.p, sizeof(int), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		// splits are granted to a prefix of the requests (isFullPool is monotone): the pool grows by the
		// requests that still fitted when their turn came
		int grown = std::min(growth, std::max(0, c->SP.limit - 4 - n));
		if (growth > 0 && grown < growth) {
			// some requests were refused: the granted prefix ends where first + 4 >= limit; recompute exactly on the host
			std::vector<int> cnt(count);
			CUDA_TRY(cudaMemcpy(cnt.data(), c->sp_count.p + prev, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost));
			int off = 0;
			for (int k = 0; k < count; k++) {
				const int ck = cnt[k] == kReshade ? 2 : cnt[k];	// granted in-place splits were re-marked by the append kernel
				if (ck == 0) continue;
				if (n + off + 4 >= c->SP.limit) break;
				off += ck;
			}
			grown = off;
		}
		const bool device_advanced = n + growth + 4 < c->SP.limit;	// split_scan_offsets_kernel already moved trinum
		prev = n;
		n += grown;
		c->sp_host_trinum = n;
		if (!device_advanced) {
			CUDA_TRY(cudaMemcpyAsync(c->sp_trinum.p, &c->sp_host_trinum, sizeof(int), cudaMemcpyHostToDevice, st));
			CUDA_TRY(cudaStreamSynchronize(st));
		}
	}
	split_finish_shot_kernel<<<1, 1, 0, st>>>(c->SP, sh);
	c->launches++;
	c->trinum = c->sp_host_trinum;
	*converged = false;
	return RADK_OK;
}

int upload_scene(radk_ctx *c, const radk_scene *s) {
	const radk_opts &o = c->opt;
	cudaStream_t st = c->stream;
	c->trinum = s->trinum;
	c->n0 = s->n0;
	const int world = std::max(1, o.world);
	FlatScene F;
	const int frc = flatten_scene(s, o.rank, world, F);
	if (frc != RADK_OK) { set_error(F.error); return frc; }
	const int R = F.R, U = F.U;
	c->R = R, c->U = U, c->leaves_total = F.leaves_total;
	c->recv_slot = F.recv_slot;
	c->recv_parent = F.recv_parent;
	c->h_slot_to_recv = F.slot_to_recv;
	CUDA_TRY(c->cell_leaf.upload(F.cell_leaf, st));
	CUDA_TRY(c->leaf_box.upload(F.leaf_box, st));
	CUDA_TRY(c->list.upload(F.list, st));
	CUDA_TRY(c->list4.upload(F.list4, st));
	CUDA_TRY(c->leaf_list4.upload(F.leaf_list4, st));
	CUDA_TRY(c->tri_leaf_begin.upload(F.tri_leaf_begin, st));
	CUDA_TRY(c->tri_leaves.upload(F.tri_leaves, st));
	CUDA_TRY(c->occ.upload(F.occ, st));
	for (int v = 0; v < 3; v++) {
		CUDA_TRY(c->geo[v].upload(F.geo[v], st));
		CUDA_TRY(c->st[v].upload(F.st[v], st));
		CUDA_TRY(c->st0[v].upload(F.st[v], st));
	}
	CUDA_TRY(c->slot_to_recv.upload(F.slot_to_recv, st));
	CUDA_TRY(c->slot.upload(F.recv_slot, st));
	CUDA_TRY(c->par_n.upload(F.par_n, st));
	CUDA_TRY(c->par_fk.upload(F.par_fk, st));
	CUDA_TRY(c->vid.upload(F.vid, st));
	CUDA_TRY(c->uv.upload(F.uv, st));
	CUDA_TRY(c->uv_leaf.upload(F.uv_leaf, st));
	CUDA_TRY(c->ffh.alloc(U));
	CUDA_TRY(c->vstat.alloc(U));
	CUDA_TRY(cudaMemsetAsync(c->ffh.p, 0, std::max(U, 1) * sizeof(float2), st));
	CUDA_TRY(cudaMemsetAsync(c->vstat.p, 0, std::max(U, 1) * sizeof(unsigned), st));

	// ---- selection / bookkeeping state
	c->max_blocks = std::max(1, (R + 255) / 256);
	CUDA_TRY(c->partial_key.alloc(c->max_blocks));
	CUDA_TRY(c->partial_cnt.alloc(c->max_blocks));
	CUDA_TRY(c->ticket.alloc(1));
	CUDA_TRY(c->bound.alloc(1));
	CUDA_TRY(c->flags.alloc(4));
	CUDA_TRY(c->shots_done.alloc(1));
	CUDA_TRY(c->shot_log.alloc(kShotLogCap));
	CUDA_TRY(c->stats.alloc(1));
	CUDA_TRY(c->rec.alloc(2 + kMaxSources));
	c->use_fast = c->use_dedup && !o.collect_stats && !o.per_shot_kernels;
	if (world > kMaxWorld) { set_error("world larger than 8"); return RADK_ERR_ARG; }
	if (c->use_fast) {
		// form-factor cache rows = candidates per round (at most kMaxRows = 128), less if the cache would take more than a
		// quarter of the free memory (RADK_FF_ROWS / opts.ff_rows: experiments and tests; every rank of a sharded run must
		// use the same value, which the candidate mail checks)
		size_t free_b = 0, total_b = 0;
		CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
		// Default candidates per round by receivers per rank: < 50 000: 64, < 200 000: 48, else 32.  Small shards are bound by
		// the per-round latency chain and by form-factor launches too small to fill the device; with long schedules more
		// rounds are cut short by an overtaking receiver and pay the apply pass twice.  Measured at 32 / 48 / 64 rows, ms to
		// convergence: church (116 791 receivers) 3.92 / 3.83 / 3.81, box (98 734) 28.2 / 29.3 / 31.7, room (332 204, first 3000
		// shots) 57.2 / 57.3 / 60.5; church on 8 GPUs (14 599 per rank) 2.07 / - / 1.33, 1.39 at 128.  The rule only uses
		// numbers every rank agrees on.
		const char *env_rows = getenv("RADK_FF_ROWS");
		// -parallel_src P: an iteration needs P rows at once; four iterations per round where kMaxRows allows it
		const int P = std::max(1, std::min(o.parallel_src, kMaxSources));
		const int def_rows = P > 1 ? std::min(kMaxRows, std::max(32, (4 * P + 31) / 32 * 32)) : (F.leaves_total / world < 50000 ? 64 : (F.leaves_total / world < 200000 ? 48 : 32));
		int rows = o.ff_rows > 0 ? std::min(o.ff_rows, kMaxRows) : (env_rows && atoi(env_rows) > 0 ? std::min(atoi(env_rows), kMaxRows) : def_rows);
		while (rows > 2 && (size_t)rows * std::max(U, 1) * sizeof(float2) > free_b / 4) rows >>= 1;
		if (rows < P) { set_error("the form-factor cache cannot hold one row per -parallel_src shooter"); return RADK_ERR_LIMIT; }
		c->rows = rows;
		CUDA_TRY(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem_bytes(kMaxRows)));
		{ const int prc = preload_round_kernels(); if (prc) return prc; }
		CUDA_TRY(c->ffcache.alloc((size_t)rows * std::max(U, 1)));
		CUDA_TRY(c->row_slot.alloc(kMaxRows));
		CUDA_TRY(c->slot_row.alloc(std::max(s->trinum, 1)));
		CUDA_TRY(c->todo.alloc(kMaxRows));
		CUDA_TRY(c->fast_ints.alloc(4));
		CUDA_TRY(c->row_rec.alloc(kMaxRows));
		CUDA_TRY(c->group_key.alloc(kPredictGroups));
		CUDA_TRY(c->fast_uints.alloc(4));
		CUDA_TRY(c->handoff.alloc(1));
		CUDA_TRY(c->sched.alloc(kMaxChain));
		for (int v = 0; v < 3; v++) CUDA_TRY(c->st_prev[v].alloc(R));
		CUDA_TRY(c->counters.alloc(8));
		CUDA_TRY(c->budget.alloc(1));
		CUDA_TRY(c->box.alloc(1));
		CUDA_TRY(c->pred_round.alloc(1));
		CUDA_TRY(c->fast_error.alloc(1));
		CUDA_TRY(c->exact.alloc(2 + kExactCap + kMaxRows));
		CUDA_TRY(c->cand_count.alloc(1));
		CUDA_TRY(c->cand_ff.alloc((size_t)kMaxRows * kMaxRows * 3));
		CUDA_TRY(cudaMemsetAsync(c->box.p, 0, sizeof(PeerBox), st));
		const int mask_words = (s->nnodes + 31) / 32;
		CUDA_TRY(c->row_mask.alloc((size_t)kMaxRows * mask_words));
		FastState &F = c->F;
		F.ff = c->ffcache.p, F.rows = rows, F.row_slot = c->row_slot.p, F.slot_row = c->slot_row.p;
		F.rec = c->row_rec.p, F.todo = c->todo.p, F.ntodo = c->fast_ints.p;
		F.nsched = c->fast_ints.p + 1, F.viol = c->fast_ints.p + 2;
		F.group_key = c->group_key.p, F.ticket = c->fast_uints.p, F.apply_ticket = c->fast_uints.p + 1, F.chain_ticket = c->fast_uints.p + 3;
		F.handoff = c->handoff.p, F.budget = c->budget.p;
		F.sched = c->sched.p, F.counters = c->counters.p;
		for (int v = 0; v < 3; v++) F.st_prev[v] = c->st_prev[v].p;
		F.row_mask = c->row_mask.p, F.mask_words = mask_words;
		for (int t = 0; t < kMaxWorld; t++) F.box[t] = nullptr;
		F.box[o.rank] = c->box.p;
		F.rank = o.rank, F.world = world;
		F.pred_round = c->pred_round.p, F.error = c->fast_error.p;
		F.parallel_src = P, F.local_count = c->fast_uints.p + 2, F.cand_count = c->cand_count.p, F.exact = c->exact.p;
		F.cand_ff = c->cand_ff.p;
		F.dbg = nullptr;
		if (getenv("RADK_DEBUG_PREDICT")) { CUDA_TRY(c->dbg.alloc(16 + 64 * 16)); CUDA_TRY(cudaMemsetAsync(c->dbg.p, 0, (16 + 64 * 16) * 8, st)); F.dbg = c->dbg.p; }
		c->peers_connected = world == 1;
		int rc2 = reset_fast_state(c);
		if (rc2) return rc2;
	}
	CUDA_TRY(cudaMemsetAsync(c->ticket.p, 0, sizeof(unsigned), st));
	CUDA_TRY(cudaMemsetAsync(c->bound.p, 0xff, sizeof(unsigned long long), st));
	CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, 4 * sizeof(int), st));
	CUDA_TRY(cudaMemsetAsync(c->shots_done.p, 0, sizeof(long long), st));
	CUDA_TRY(cudaMemsetAsync(c->stats.p, 0, sizeof(DevStats), st));
	CUDA_TRY(cudaMemsetAsync(c->rec.p, 0, (2 + kMaxSources) * sizeof(ShooterRec), st));

	DevScene &S = c->S;
	S.tree.cell_leaf = c->cell_leaf.p;
	S.tree.leaf_box = c->leaf_box.p;
	S.tree.list = c->list.p;
	S.tree.list4 = c->list4.p, S.tree.leaf_list4 = c->leaf_list4.p;
	S.tree.tri_leaf_begin = c->tri_leaf_begin.p, S.tree.tri_leaves = c->tri_leaves.p;
	S.tree.nnodes = s->nnodes;
	S.tree.occ = c->occ.p, S.tree.nocc = s->n0;
	for (int k = 0; k < 3; k++) S.tree.G0[k] = s->G0[k], S.tree.G1[k] = s->G1[k];
	S.tree.min_grid_len = s->min_grid_len;
	S.tree.grid_num = s->grid_num;
	S.tree.inv_min_grid_len = 1.0 / (double)s->min_grid_len;
	S.tree.inv_min_grid_len_f = (float)S.tree.inv_min_grid_len;
	S.tree.cell_eps = 4e-7f * (float)(s->grid_num + 1) + 1e-4f;
	S.R = R, S.U = U, S.n0 = s->n0;
	for (int v = 0; v < 3; v++) S.geo[v] = c->geo[v].p, S.st[v] = c->st[v].p;
	S.vid = c->vid.p, S.slot = c->slot.p, S.uv = c->uv.p, S.uv_leaf = c->uv_leaf.p, S.ffh = c->ffh.p, S.vstat = c->vstat.p;
	S.par_n = c->par_n.p, S.par_fk = c->par_fk.p, S.slot_to_recv = c->slot_to_recv.p;
	S.converge = o.converge;
	S.parallel_src = std::max(1, std::min(o.parallel_src, kMaxSources));
	S.stats = c->stats.p;
	S.sel.partial_key = c->partial_key.p, S.sel.partial_cnt = c->partial_cnt.p, S.sel.ticket = c->ticket.p;
	S.sel.bound = c->bound.p;
	S.sel.ncand = c->flags.p, S.sel.nsrc = c->flags.p + 1, S.sel.converged = c->flags.p + 2;
	S.sel.shots_done = c->shots_done.p;
	S.sel.shot_log = c->shot_log.p, S.sel.shot_log_cap = kShotLogCap;
	CUDA_TRY(cudaStreamSynchronize(st));
	CUDA_TRY(cudaGetLastError());
	c->have_current = false;
	c->cur = 0;
	return RADK_OK;
}

inline int blocks_for(int n, int bs) { return std::max(1, (n + bs - 1) / bs); }

// K1 stand-alone: (re)select the current shooter into rec[cur]
void launch_select(radk_ctx *c, ShooterRec *out, int which) {
	select_kernel<<<std::min(c->max_blocks, 4 * 148), 256, 0, c->stream>>>(c->S, out, which);
	c->launches++;
}

void launch_single_source_shot(radk_ctx *c) {
	const bool stats = c->opt.collect_stats != 0;
	ShooterRec *cur = c->rec.p + c->cur, *nxt = c->rec.p + (c->cur ^ 1);
	const int nb_v = blocks_for(2 * c->U, 128);
	if (stats) vertex_ff_kernel<true><<<nb_v, 128, 0, c->stream>>>(c->S, cur);
	else vertex_ff_kernel<false><<<nb_v, 128, 0, c->stream>>>(c->S, cur);
	const int nb_u = blocks_for(c->R, 256);
	if (stats) update_select_kernel<true><<<nb_u, 256, 0, c->stream>>>(c->S, cur, nxt);
	else update_select_kernel<false><<<nb_u, 256, 0, c->stream>>>(c->S, cur, nxt);
	c->cur ^= 1;
	c->launches += 2;
}

void launch_multi_source_iteration(radk_ctx *c) {
	const bool stats = c->opt.collect_stats != 0;
	ShooterRec *list = c->rec.p + 2;
	const int P = c->S.parallel_src;
	for (int j = 0; j < P; j++) launch_select(c, list + j, j);
	const int nb = blocks_for(c->R, 128);
	if (stats) shade_receivers_kernel<true><<<nb, 128, 0, c->stream>>>(c->S, list, P);
	else shade_receivers_kernel<false><<<nb, 128, 0, c->stream>>>(c->S, list, P);
	after_generic_pass_kernel<<<1, 1, 0, c->stream>>>(c->S, list);
	c->launches += 2;
}

int read_flags(radk_ctx *c) {
	CUDA_TRY(cudaMemcpyAsync(c->h_flags, c->flags.p, sizeof(*c->h_flags), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaMemcpyAsync(c->h_shots, c->shots_done.p, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
	if (c->use_fast) {
		CUDA_TRY(cudaMemcpyAsync(c->h_budget, c->budget.p, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaMemcpyAsync(c->h_error, c->fast_error.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	}
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	if (c->use_fast && *c->h_error == 2) { set_error("-parallel_src selection: more than 2048 receivers above the selection threshold share fewer than P selection groups"); return RADK_ERR_LIMIT; }
	if (c->use_fast && *c->h_error) {
		set_error("a peer rank did not answer within the device-side timeout (exchange " + std::to_string(*c->h_error) + ": 11 candidates, 12 candidate records, 13 first violation)");
		return RADK_ERR_STATE;
	}
	return RADK_OK;
}

// RADK_DEBUG_PREDICT: mean time (us, after the round's predict start) of the round's events over the last <= 64 rounds;
// out[0..10] = the events of dbg_trace (fast_kernels.cuh), out[11] = median gap from the commit's bookkeeping to the next
// round's start, out[12] = rounds averaged, out[13] = mismatches of the top-64 selection self-check.  -1: not recorded.
int round_timeline(radk_ctx *c, double out[14]) {
	for (int k = 0; k < 14; k++) out[k] = -1.0;
	if (!c->dbg.p) return RADK_OK;
	std::vector<unsigned long long> t(16 + 64 * 16);
	CUDA_TRY(cudaMemcpy(t.data(), c->dbg.p, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	const int nr = (int)std::min<unsigned long long>(t[15], 64);
	double sum[11] = {0};
	int cnt[11] = {0};
	std::vector<double> gaps;
	for (int r = 0; r < nr; r++) {
		const unsigned long long *q = t.data() + 16 + 16 * r;
		for (int k = 0; k < 11; k++)
			if (q[0] && q[k] >= q[0] && q[k] - q[0] < 100000000ull) sum[k] += (double)(q[k] - q[0]), cnt[k]++;
		const unsigned long long *nx = t.data() + 16 + 16 * ((r + 1) % 64);
		if (nx[0] > q[10] && nx[0] - q[10] < 1000000ull) gaps.push_back((double)(nx[0] - q[10]));
	}
	for (int k = 0; k < 11; k++)
		if (cnt[k]) out[k] = sum[k] / cnt[k] / 1e3;
	if (!gaps.empty()) {	// the median: the gap between two steps (host work) is not a round-to-round gap
		std::sort(gaps.begin(), gaps.end());
		out[11] = gaps[gaps.size() / 2] / 1e3;
	}
	out[12] = nr, out[13] = (double)t[14];
	return RADK_OK;
}

}  // namespace

extern "C" {

int radk_ctx_create(const radk_scene *s, const radk_opts *o, radk_ctx **out) {
	if (!s || !out || !s->tri || !s->nodes || !s->list) { set_error("radk_ctx_create: null argument"); return RADK_ERR_ARG; }
	radk_opts opt;
	if (o) opt = *o; else radk_default_opts(&opt);
	if (opt.world <= 0) opt.world = 1;
	if (opt.rank < 0 || opt.rank >= opt.world) { set_error("radk_ctx_create: rank outside [0, world)"); return RADK_ERR_ARG; }
	if (opt.split_enabled && opt.world > 1) { set_error("radk_ctx_create: adaptive splitting runs on one GPU"); return RADK_ERR_ARG; }
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		set_error(std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
		return RADK_ERR_CUDA;
	}
	CUDA_TRY(cudaSetDevice(opt.device));
	radk_ctx *c = new radk_ctx();
	c->opt = opt;
	c->device = opt.device;
	// vertex sharing + form-factor cache: single-source always, -parallel_src > 1 on the batched path only (statistics
	// mode and per-shot kernels take the generic per-receiver kernel)
	c->use_dedup = opt.dedup >= 0 && (opt.parallel_src <= 1 || (!opt.collect_stats && !opt.per_shot_kernels));
	int rc = RADK_OK;
	do {
		if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
		    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
		    cudaMallocHost((void **)&c->h_flags, sizeof(*c->h_flags)) != cudaSuccess ||
		    cudaMallocHost((void **)&c->h_shots, sizeof(long long)) != cudaSuccess ||
		    cudaMallocHost((void **)&c->h_budget, sizeof(long long)) != cudaSuccess ||
		    cudaMallocHost((void **)&c->h_error, sizeof(int)) != cudaSuccess) {
			set_error(std::string("context setup failed: ") + cudaGetErrorString(cudaGetLastError()));
			rc = RADK_ERR_CUDA;
			break;
		}
		rc = opt.split_enabled ? upload_split_scene(c, s) : upload_scene(c, s);
	} while (0);
	if (rc != RADK_OK) { radk_ctx_destroy(c); return rc; }
	*out = c;
	return RADK_OK;
}

void radk_ctx_destroy(radk_ctx *c) {
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->dbg.p) {	// RADK_DEBUG_PREDICT: phases of the last predict_kernel launch
		unsigned long long t[16 + 64 * 16];
		if (cudaMemcpy(t, c->dbg.p, sizeof(t), cudaMemcpyDeviceToHost) == cudaSuccess) {
			fprintf(stderr, "[radk] predict phases (ns): scan+ticket %llu, top-K %llu, cache rebuild %llu, shooter records %llu, leaf masks %llu\n",
				t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4]);
			static const char *ev[11] = {"predict", "predict_end", "merge", "merge_end", "ff_cand", "ff_batch", "chain", "chain_end", "apply", "commit", "commit_end"};
			double tl[14];
			if (round_timeline(c, tl) == RADK_OK) {
				fprintf(stderr, "[radk] round timeline, mean us after the round's predict start (rank %d, %d rounds):", c->opt.rank, (int)tl[12]);
				for (int k = 0; k < 11; k++)
					if (tl[k] >= 0) fprintf(stderr, " %s %.1f", ev[k], tl[k]);
				fprintf(stderr, " | gap to next round %.1f\n", tl[11]);
			}
			const unsigned long long rounds = t[15];
			fprintf(stderr, "[radk] top-64 selection self-check: %llu mismatches against the full sort in %llu rounds\n", t[14], rounds);
		}
	}
	if (c->h_flags) cudaFreeHost(c->h_flags);
	if (c->h_shots) cudaFreeHost(c->h_shots);
	if (c->h_stage) cudaFreeHost(c->h_stage);
	for (auto e : c->stage_ev) if (e) cudaEventDestroy(e);
	if (c->h_budget) cudaFreeHost(c->h_budget);
	if (c->h_error) cudaFreeHost(c->h_error);
	if (c->loop_exec) cudaGraphExecDestroy(c->loop_exec);
	if (c->loop_graph) cudaGraphDestroy(c->loop_graph);
	for (void *q : c->ipc_opened) cudaIpcCloseMemHandle(q);
	if (c->ev0) cudaEventDestroy(c->ev0);
	if (c->ev1) cudaEventDestroy(c->ev1);
	if (c->ev_fork) cudaEventDestroy(c->ev_fork);
	if (c->ev_join) cudaEventDestroy(c->ev_join);
	if (c->side) cudaStreamDestroy(c->side);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

int radk_ctx_reset(radk_ctx *c, const radk_scene *s) {
	if (c && c->split_mode) { set_error("radk_ctx_reset: not available on adaptive-split contexts"); return RADK_ERR_STATE; }
	if (!c || !s) { set_error("radk_ctx_reset: null argument"); return RADK_ERR_ARG; }
	if (s->trinum != c->trinum) { set_error("radk_ctx_reset: scene size changed"); return RADK_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	if (!c->h_stage) CUDA_TRY(cudaMallocHost((void **)&c->h_stage, 3 * (size_t)std::max(c->R, 1) * sizeof(float4)));
	const size_t R = (size_t)c->R;
	// gather from the caller's TriStore image into pinned planes, one receiver range at a time: the copy of a
	// range is in flight while the next one is gathered
	for (int k = 0; k < kStageChunks; k++) {
		const size_t b = R * k / kStageChunks, e = R * (k + 1) / kStageChunks;
		if (e == b) continue;
		host_parallel_for(e - b, [=](size_t r0, size_t r1) {
			for (size_t r = b + r0; r < b + r1; r++) {
				// the records are 236 bytes apart in slot order, the receivers are in Morton order: prefetch ahead
				if (r + 12 < b + r1) {
					const radk_triangle &n = s->tri[c->recv_slot[r + 12]];
					__builtin_prefetch(&n.accB[0][0]);
					__builtin_prefetch(&n.deltaB[2]);
				}
				const radk_triangle &t = s->tri[c->recv_slot[r]];
				for (int v = 0; v < 3; v++) c->h_stage[v * R + r] = make_float4(t.accB[v][0], t.accB[v][1], t.accB[v][2], t.deltaB[v]);
			}
		});
		for (int v = 0; v < 3; v++)
			CUDA_TRY(cudaMemcpyAsync(c->st[v].p + b, c->h_stage + v * R + b, (e - b) * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
	}
	for (int v = 0; v < 3; v++)
		CUDA_TRY(cudaMemcpyAsync(c->st0[v].p, c->st[v].p, R * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
	CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, 4 * sizeof(int), c->stream));
	CUDA_TRY(cudaMemsetAsync(c->shots_done.p, 0, sizeof(long long), c->stream));
	CUDA_TRY(cudaMemsetAsync(c->stats.p, 0, sizeof(DevStats), c->stream));
	{ int rc2 = reset_fast_state(c); if (rc2) return rc2; }
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	c->have_current = false;
	c->device_ms = 0;
	return RADK_OK;
}

int radk_ctx_rewind(radk_ctx *c) {
	if (c && c->split_mode) { set_error("radk_ctx_rewind: not available on adaptive-split contexts"); return RADK_ERR_STATE; }
	if (!c) { set_error("radk_ctx_rewind: null context"); return RADK_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	for (int v = 0; v < 3; v++)
		CUDA_TRY(cudaMemcpyAsync(c->st[v].p, c->st0[v].p, (size_t)c->R * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
	CUDA_TRY(cudaMemsetAsync(c->flags.p, 0, 4 * sizeof(int), c->stream));
	CUDA_TRY(cudaMemsetAsync(c->shots_done.p, 0, sizeof(long long), c->stream));
	{ int rc2 = reset_fast_state(c); if (rc2) return rc2; }	// cached form factors are dropped: a re-run recomputes everything
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	c->have_current = false;
	return RADK_OK;
}

int radk_ctx_sizes(const radk_ctx *c, int *receivers, int *unique_vertices, int *occluders, int *nodes) {
	if (!c) { set_error("radk_ctx_sizes: null context"); return RADK_ERR_ARG; }
	if (receivers) *receivers = c->R;
	if (unique_vertices) *unique_vertices = c->U;
	if (occluders) *occluders = c->n0;
	if (nodes) *nodes = (int)(c->leaf_box.n / 2);
	return RADK_OK;
}

int radk_ctx_profile(radk_ctx *c, long long max_shots, double ms[3], long long n[3], long long *shots) {
	if (c && c->split_mode) { set_error("radk_ctx_profile: not available on adaptive-split contexts"); return RADK_ERR_STATE; }
	if (!c || !ms || !n) { set_error("radk_ctx_profile: bad argument"); return RADK_ERR_ARG; }
	if (!c->use_dedup) { set_error("radk_ctx_profile: single-source contexts only"); return RADK_ERR_STATE; }
	CUDA_TRY(cudaSetDevice(c->device));
	int rc = read_flags(c);
	if (rc) return rc;
	const long long before = *c->h_shots;
	for (int k = 0; k < 3; k++) ms[k] = 0, n[k] = 0;
	constexpr int kProfRounds = 8;
	struct Events {		// destroyed on every exit, also the error returns below
		cudaEvent_t e[4 * kProfRounds] = {};
		radk_ctx *c;
		bool ok = false;
		explicit Events(radk_ctx *ctx) : c(ctx) {}
		~Events() {
			for (auto &x : e) if (x) cudaEventDestroy(x);
			if (!ok) c->have_current = false;	// a partial iteration leaves rec[cur] behind the device state
		}
	} evh(c);
	cudaEvent_t *ev = evh.e;
	for (int i = 0; i < 4 * kProfRounds; i++) CUDA_TRY(cudaEventCreate(&ev[i]));
	const bool stats = c->opt.collect_stats != 0;
	long long done = 0;
	if (c->use_fast) {
		*c->h_budget = max_shots < 0 ? (1ll << 60) : max_shots;
		CUDA_TRY(cudaMemcpyAsync(c->budget.p, c->h_budget, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		while (true) {
			// several rounds are enqueued back to back so the GPU never waits for the host between kernels
			for (int i = 0; i < kProfRounds; i++)
				if ((rc = launch_fast_round(c, ev + 4 * i))) return rc;
			if ((rc = read_flags(c))) return rc;
			for (int i = 0; i < kProfRounds; i++)
				for (int k = 0; k < 3; k++) {
					float t = 0;
					CUDA_TRY(cudaEventElapsedTime(&t, ev[4 * i + k], ev[4 * i + k + 1]));
					ms[k] += t, n[k]++;
				}
			done = *c->h_shots - before;
			if (c->h_flags->converged || *c->h_budget <= 0) break;
		}
	} else {
		if (!c->have_current) {
			CUDA_TRY(cudaEventRecord(ev[0], c->stream));
			launch_select(c, c->rec.p + c->cur, 0);
			CUDA_TRY(cudaEventRecord(ev[1], c->stream));
			CUDA_TRY(cudaEventSynchronize(ev[1]));
			float t = 0;
			CUDA_TRY(cudaEventElapsedTime(&t, ev[0], ev[1]));
			ms[0] += t, n[0]++;
			c->have_current = true;
		}
		while (max_shots < 0 || done < max_shots) {
			ShooterRec *cur = c->rec.p + c->cur, *nxt = c->rec.p + (c->cur ^ 1);
			CUDA_TRY(cudaEventRecord(ev[0], c->stream));
			if (stats) vertex_ff_kernel<true><<<blocks_for(2 * c->U, 128), 128, 0, c->stream>>>(c->S, cur);
			else vertex_ff_kernel<false><<<blocks_for(2 * c->U, 128), 128, 0, c->stream>>>(c->S, cur);
			CUDA_TRY(cudaEventRecord(ev[1], c->stream));
			if (stats) update_select_kernel<true><<<blocks_for(c->R, 256), 256, 0, c->stream>>>(c->S, cur, nxt);
			else update_select_kernel<false><<<blocks_for(c->R, 256), 256, 0, c->stream>>>(c->S, cur, nxt);
			CUDA_TRY(cudaEventRecord(ev[2], c->stream));
			c->cur ^= 1;
			c->launches += 2;
			if ((rc = read_flags(c))) return rc;
			if (*c->h_shots - before == done) break;	// nothing ran: converged
			done = *c->h_shots - before;
			float t1 = 0, t2 = 0;
			CUDA_TRY(cudaEventElapsedTime(&t1, ev[0], ev[1]));
			CUDA_TRY(cudaEventElapsedTime(&t2, ev[1], ev[2]));
			ms[1] += t1, n[1]++, ms[2] += t2, n[2]++;
			if (c->h_flags->converged) break;
		}
	}
	evh.ok = true;
	if (shots) *shots = done;
	CUDA_TRY(cudaGetLastError());
	return RADK_OK;
}

int radk_ctx_step(radk_ctx *c, long long max_shots, long long *shots_done, int *converged) {
	if (!c) { set_error("radk_ctx_step: null context"); return RADK_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	if (c->opt.world > 1 && !c->use_fast) { set_error("multi-GPU contexts support the default (batched) path only"); return RADK_ERR_STATE; }
	if (c->opt.world > 1 && !c->peers_connected) { set_error("radk_ctx_step: call radk_ctx_peer_connect first"); return RADK_ERR_STATE; }
	long long before = 0;
	if (c->h_error) *c->h_error = 0;
	int rc = read_flags(c);
	if (rc) return rc;
	before = *c->h_shots;
	long long done = 0;
	int conv = 0;
	CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
	if (c->split_mode) {
		while (max_shots < 0 || done < max_shots) {
			bool cv = false;
			if ((rc = split_one_shot(c, &cv))) return rc;
			if (cv) { conv = 1; break; }
			done++;
		}
	} else if (c->use_fast) {
		*c->h_budget = max_shots < 0 ? (1ll << 60) : max_shots;
		const long long budget0 = *c->h_budget;
		CUDA_TRY(cudaMemcpyAsync(c->budget.p, c->h_budget, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		if (budget0 > 0 && build_loop_graph(c)) {
			// the whole step is ONE launch: the rounds repeat on the device until converged / out of budget
			unsigned long long rounds0 = 0, rounds1 = 0;
			CUDA_TRY(cudaMemcpyAsync(&rounds0, c->counters.p + 1, sizeof(rounds0), cudaMemcpyDeviceToHost, c->stream));
			CUDA_TRY(cudaGraphLaunch(c->loop_exec, c->stream));
			CUDA_TRY(cudaMemcpyAsync(&rounds1, c->counters.p + 1, sizeof(rounds1), cudaMemcpyDeviceToHost, c->stream));
			if ((rc = read_flags(c))) return rc;
			c->launches += (long long)(rounds1 - rounds0) * c->launches_per_round;
			done = *c->h_shots - before;
			conv = c->h_flags->converged ? 1 : 0;
		} else
		while (budget0 > 0) {
			// a round runs ~rows shots; enqueue a few rounds per host round trip
			const long long left = budget0 - done;
			const int rounds = (int)std::max<long long>(1, std::min<long long>(4, (left + c->rows - 1) / c->rows));
			for (int i = 0; i < rounds; i++)
				if ((rc = launch_fast_round(c, nullptr))) return rc;
			if ((rc = read_flags(c))) return rc;
			done = *c->h_shots - before;
			if (c->h_flags->converged) { conv = 1; break; }
			if (*c->h_budget <= 0) break;
		}
	} else if (c->use_dedup) {
		if (!c->have_current) { launch_select(c, c->rec.p + c->cur, 0); c->have_current = true; }
		while (max_shots < 0 || done < max_shots) {
			long long batch = kBatchShots;
			if (max_shots >= 0) batch = std::min<long long>(batch, max_shots - done);
			for (long long i = 0; i < batch; i++) launch_single_source_shot(c);
			if ((rc = read_flags(c))) return rc;
			done = *c->h_shots - before;
			if (c->h_flags->converged) { conv = 1; break; }
		}
	} else {
		while (max_shots < 0 || done < max_shots) {
			long long batch = std::min<long long>(kBatchShots, max_shots < 0 ? kBatchShots : max_shots - done);
			for (long long i = 0; i < batch; i++) launch_multi_source_iteration(c);
			if ((rc = read_flags(c))) return rc;
			done = *c->h_shots - before;
			if (c->h_flags->converged) { conv = 1; break; }
		}
		c->have_current = false;
	}
	CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
	CUDA_TRY(cudaEventSynchronize(c->ev1));
	float ms = 0;
	CUDA_TRY(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
	c->device_ms += ms;
	CUDA_TRY(cudaGetLastError());
	if (shots_done) *shots_done = done;
	if (converged) *converged = conv;
	return RADK_OK;
}

int radk_ctx_select(radk_ctx *c, int *slots, int cap, int *n) {
	if (c && c->split_mode) { set_error("radk_ctx_select: not available on adaptive-split contexts"); return RADK_ERR_STATE; }
	if (!c || !n) { set_error("radk_ctx_select: null argument"); return RADK_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	ShooterRec *list = c->rec.p + 2;
	const int P = c->S.parallel_src;
	for (int j = 0; j < P; j++) launch_select(c, list + j, j);
	int rc = read_flags(c);
	if (rc) return rc;
	std::vector<ShooterRec> h(P);
	CUDA_TRY(cudaMemcpy(h.data(), list, P * sizeof(ShooterRec), cudaMemcpyDeviceToHost));
	int k = 0;
	for (int j = 0; j < c->h_flags->nsrc && j < P; j++)
		if (h[j].valid && k < cap && slots) slots[k++] = h[j].slot;
	*n = std::min(c->h_flags->nsrc, P);
	c->have_current = false;	// flags.converged was rewritten by the stand-alone selection
	return RADK_OK;
}

int radk_ctx_shooter_log(radk_ctx *c, int *slots, int cap, long long *n) {
	if (!c || !n) { set_error("radk_ctx_shooter_log: null argument"); return RADK_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	int rc = read_flags(c);
	if (rc) return rc;
	*n = *c->h_shots;
	const long long have = std::min<long long>(*n, kShotLogCap);
	const long long want = std::min<long long>(have, cap);
	if (slots && want > 0) {
		std::vector<int> ring(kShotLogCap);
		CUDA_TRY(cudaMemcpy(ring.data(), c->shot_log.p, sizeof(int) * (size_t)std::min<long long>(*n, kShotLogCap), cudaMemcpyDeviceToHost));
		for (long long i = 0; i < want; i++) slots[i] = ring[(size_t)((*n - want + i) % kShotLogCap)];
	}
	return RADK_OK;
}

int radk_ctx_trinum(const radk_ctx *c) { return c ? c->trinum : -1; }

int radk_ctx_download(radk_ctx *c, float *accB, float *deltaB, int *parent, int trinum) {
	if (!c) { set_error("radk_ctx_download: null context"); return RADK_ERR_ARG; }
	if (trinum < c->trinum) { set_error("radk_ctx_download: buffers smaller than trinum"); return RADK_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	if (c->split_mode) {
		const int n = c->trinum;
		std::vector<float4> hs[3], hp(n);
		for (int v = 0; v < 3; v++) {
			hs[v].resize(n);
			CUDA_TRY(cudaMemcpyAsync(hs[v].data(), c->sp_st[v].p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
		}
		CUDA_TRY(cudaMemcpyAsync(hp.data(), c->sp_p[0].p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		for (int i = 0; i < n; i++) {
			int par;
			memcpy(&par, &hp[i].w, 4);
			if (parent) parent[i] = par;
			if (par < 0) continue;
			for (int v = 0; v < 3; v++) {
				if (accB) accB[(size_t)i * 9 + v * 3] = hs[v][i].x, accB[(size_t)i * 9 + v * 3 + 1] = hs[v][i].y, accB[(size_t)i * 9 + v * 3 + 2] = hs[v][i].z;
				if (deltaB) deltaB[(size_t)i * 3 + v] = hs[v][i].w;
			}
		}
		return RADK_OK;
	}
	if (!c->h_stage) CUDA_TRY(cudaMallocHost((void **)&c->h_stage, 3 * (size_t)std::max(c->R, 1) * sizeof(float4)));
	const size_t R = (size_t)c->R;
	// one receiver range at a time: a range is scattered into the caller's slot-indexed buffers while the
	// copies of the following ranges are still in flight
	if (!c->stage_ev[0])
		for (auto &e : c->stage_ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	for (int k = 0; k < kStageChunks; k++) {
		const size_t b = R * k / kStageChunks, e = R * (k + 1) / kStageChunks;
		for (int v = 0; v < 3 && e > b; v++)
			CUDA_TRY(cudaMemcpyAsync(c->h_stage + v * R + b, c->st[v].p + b, (e - b) * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaEventRecord(c->stage_ev[k], c->stream));
	}
	for (int k = 0; k < kStageChunks; k++) {
		const size_t b = R * k / kStageChunks, e = R * (k + 1) / kStageChunks;
		CUDA_TRY(cudaEventSynchronize(c->stage_ev[k]));
		if (e == b) continue;
		host_parallel_for(e - b, [=](size_t r0, size_t r1) {
			for (size_t r = b + r0; r < b + r1; r++) {
				if (r + 12 < b + r1) {	// scattered stores: fetch the target lines for writing ahead of time
					const size_t ns = (size_t)c->recv_slot[r + 12];
					if (accB) { __builtin_prefetch(accB + ns * 9, 1); __builtin_prefetch(accB + ns * 9 + 8, 1); }
					if (deltaB) __builtin_prefetch(deltaB + ns * 3, 1);
					if (parent) __builtin_prefetch(parent + ns, 1);
				}
				const size_t slot = (size_t)c->recv_slot[r];
				const float4 a0 = c->h_stage[r], a1 = c->h_stage[R + r], a2 = c->h_stage[2 * R + r];
				if (accB) {
					float *a = accB + slot * 9;
					a[0] = a0.x, a[1] = a0.y, a[2] = a0.z, a[3] = a1.x, a[4] = a1.y, a[5] = a1.z, a[6] = a2.x, a[7] = a2.y, a[8] = a2.z;
				}
				if (deltaB) deltaB[slot * 3] = a0.w, deltaB[slot * 3 + 1] = a1.w, deltaB[slot * 3 + 2] = a2.w;
				if (parent) parent[slot] = c->recv_parent[r];
			}
		});
	}
	return RADK_OK;
}

int radk_ctx_download_triangles(radk_ctx *c, radk_triangle *tri, int cap, int *trinum) {
	if (!c || !trinum) { set_error("radk_ctx_download_triangles: null argument"); return RADK_ERR_ARG; }
	if (!c->split_mode) { set_error("radk_ctx_download_triangles: only meaningful for adaptive-split contexts"); return RADK_ERR_STATE; }
	*trinum = c->trinum;
	if (!tri) return RADK_OK;
	if (cap < c->trinum) { set_error("radk_ctx_download_triangles: buffer smaller than trinum"); return RADK_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	const int n = c->trinum;
	std::vector<float4> hp[3], hs[3], hn(n);
	for (int v = 0; v < 3; v++) {
		hp[v].resize(n), hs[v].resize(n);
		CUDA_TRY(cudaMemcpyAsync(hp[v].data(), c->sp_p[v].p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaMemcpyAsync(hs[v].data(), c->sp_st[v].p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
	}
	CUDA_TRY(cudaMemcpyAsync(hn.data(), c->sp_n.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaMemcpyAsync(tri, c->sp_rec.p, (size_t)n * sizeof(radk_triangle), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	for (int i = 0; i < n; i++) {
		radk_triangle &t = tri[i];
		for (int v = 0; v < 3; v++) {
			t.p[v][0] = hp[v][i].x, t.p[v][1] = hp[v][i].y, t.p[v][2] = hp[v][i].z;
			t.accB[v][0] = hs[v][i].x, t.accB[v][1] = hs[v][i].y, t.accB[v][2] = hs[v][i].z;
			t.deltaB[v] = hs[v][i].w;
		}
		for (int k = 0; k < 3; k++) t.c[k] = (float)((double)(t.p[0][k] + t.p[1][k] + t.p[2][k]) / 3.0);
		t.n[0] = hn[i].x, t.n[1] = hn[i].y, t.n[2] = hn[i].z;
		t.area = hp[1][i].w;
		memcpy(&t.parent, &hp[0][i].w, 4);
	}
	return RADK_OK;
}

int radk_ctx_snapshot_colors(radk_ctx *c, int *rgb, int trinum) {
	if (c && c->split_mode) { set_error("radk_ctx_snapshot_colors: not available on adaptive-split contexts"); return RADK_ERR_STATE; }
	if (!c || !rgb) { set_error("radk_ctx_snapshot_colors: null argument"); return RADK_ERR_ARG; }
	if (trinum < c->trinum) { set_error("radk_ctx_snapshot_colors: buffer smaller than trinum"); return RADK_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	DevBuf<int> d;
	CUDA_TRY(d.alloc((size_t)c->R * 3));
	snapshot_kernel<<<blocks_for(c->R, 256), 256, 0, c->stream>>>(c->S, d.p);
	c->launches++;
	std::vector<int> h((size_t)c->R * 3);
	CUDA_TRY(cudaMemcpyAsync(h.data(), d.p, h.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	for (int i = 0; i < c->trinum; i++) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = -1;
	for (int r = 0; r < c->R; r++)
		for (int j = 0; j < 3; j++) rgb[(size_t)c->recv_slot[r] * 3 + j] = h[(size_t)r * 3 + j];
	return RADK_OK;
}

int radk_ctx_stats(radk_ctx *c, radk_stats *out) {
	if (!c || !out) { set_error("radk_ctx_stats: null argument"); return RADK_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	DevStats h;
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	CUDA_TRY(cudaMemcpy(&h, c->stats.p, sizeof(h), cudaMemcpyDeviceToHost));
	memset(out, 0, sizeof(*out));
	out->shots = (long long)h.shots;
	out->rays = (long long)h.rays, out->visible = (long long)h.visible, out->calff = (long long)h.calff;
	out->traced_rays = (long long)h.traced;
	out->leaf_visits = (long long)h.leaf_visits, out->descent_steps = (long long)h.descent_steps;
	out->tri_tests = (long long)h.tri_tests, out->tri_accepts = (long long)h.tri_accepts;
	out->precision_exceeded = (long long)h.precision_exceeded;
	out->receivers = (long long)h.receivers, out->splits = (long long)h.splits;
	out->kernel_launches = c->launches;
	out->ray_guard_trips = (long long)h.guard_trips;
	if (c->use_fast) {
		unsigned long long cnt[8];
		CUDA_TRY(cudaMemcpy(cnt, c->counters.p, sizeof(cnt), cudaMemcpyDeviceToHost));
		out->ff_rows_filled = (long long)cnt[0], out->rounds = (long long)cnt[1];
		out->shots_scheduled = (long long)cnt[2], out->rounds_reapplied = (long long)cnt[4];
	}
	return RADK_OK;
}

double radk_ctx_device_ms(const radk_ctx *c) { return c ? c->device_ms : 0.0; }

int radk_ctx_round_timeline(radk_ctx *c, double out[14]) {
	if (!c || !out) { set_error("radk_ctx_round_timeline: null argument"); return RADK_ERR_ARG; }
	CUDA_TRY(cudaSetDevice(c->device));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	return round_timeline(c, out);
}

int radk_ctx_shoot(radk_ctx *c, const int *slots, int n) {
	if (c && c->split_mode) { set_error("radk_ctx_shoot: not available on adaptive-split contexts"); return RADK_ERR_STATE; }
	if (!c || !slots || n <= 0 || n > kMaxSources) { set_error("radk_ctx_shoot: bad argument"); return RADK_ERR_ARG; }
	if (c->opt.world > 1) { set_error("radk_ctx_shoot: single-GPU contexts only"); return RADK_ERR_STATE; }
	CUDA_TRY(cudaSetDevice(c->device));
	std::vector<int> recv(n);
	for (int j = 0; j < n; j++) {
		if (slots[j] < 0 || slots[j] >= c->trinum || c->h_slot_to_recv[slots[j]] < 0) {
			set_error("radk_ctx_shoot: slot is not a receiver (logical or out of range)");
			return RADK_ERR_ARG;
		}
		recv[j] = c->h_slot_to_recv[slots[j]];
	}
	DevBuf<int> d;
	CUDA_TRY(d.upload(recv, c->stream));
	ShooterRec *list = c->rec.p + 2;
	make_sources_kernel<<<1, kMaxSources, 0, c->stream>>>(c->S, d.p, n, list, 1);
	const bool stats = c->opt.collect_stats != 0;
	const int nb = blocks_for(c->R, 128);
	if (stats) shade_receivers_kernel<true><<<nb, 128, 0, c->stream>>>(c->S, list, n);
	else shade_receivers_kernel<false><<<nb, 128, 0, c->stream>>>(c->S, list, n);
	after_generic_pass_kernel<<<1, 1, 0, c->stream>>>(c->S, list);
	c->launches += 3;
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	CUDA_TRY(cudaGetLastError());
	c->have_current = false;
	return RADK_OK;
}

int radk_ray_hitted(radk_ctx *c, const float *p, const float *v, const int *from, int n, int *hit) {
	if (!c || !p || !v || !from || !hit || n < 0) { set_error("radk_ray_hitted: bad argument"); return RADK_ERR_ARG; }
	if (n == 0) return RADK_OK;
	CUDA_TRY(cudaSetDevice(c->device));
	DevBuf<float> dp, dv;
	DevBuf<int> df, dh;
	CUDA_TRY(dp.alloc((size_t)n * 3));
	CUDA_TRY(dv.alloc((size_t)n * 3));
	CUDA_TRY(df.alloc(n));
	CUDA_TRY(dh.alloc(n));
	CUDA_TRY(cudaMemcpyAsync(dp.p, p, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemcpyAsync(dv.p, v, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemcpyAsync(df.p, from, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
	ray_batch_kernel<<<blocks_for(n, 128), 128, 0, c->stream>>>(c->split_mode ? c->SP.tree : c->S.tree, dp.p, dv.p, df.p, n, dh.p);
	c->launches++;
	CUDA_TRY(cudaMemcpyAsync(hit, dh.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	CUDA_TRY(cudaGetLastError());
	return RADK_OK;
}

int radk_ray_sees(radk_ctx *c, const float *p, const float *v, const int *from, int logsrc, int n, int *sees) {
	if (!c || !p || !v || !from || !sees || n < 0) { set_error("radk_ray_sees: bad argument"); return RADK_ERR_ARG; }
	if (logsrc < 0 || logsrc >= c->n0) { set_error("radk_ray_sees: logsrc is not an input triangle"); return RADK_ERR_ARG; }
	if (n == 0) return RADK_OK;
	CUDA_TRY(cudaSetDevice(c->device));
	const DevTree &T = c->split_mode ? c->SP.tree : c->S.tree;
	const size_t words = (size_t)(T.nnodes + 31) / 32;
	DevBuf<float> dp, dv;
	DevBuf<int> df, dh;
	DevBuf<unsigned> dm;
	CUDA_TRY(dp.alloc((size_t)n * 3));
	CUDA_TRY(dv.alloc((size_t)n * 3));
	CUDA_TRY(df.alloc(n));
	CUDA_TRY(dh.alloc(n));
	CUDA_TRY(dm.alloc(words));
	CUDA_TRY(cudaMemsetAsync(dm.p, 0, words * sizeof(unsigned), c->stream));
	CUDA_TRY(cudaMemcpyAsync(dp.p, p, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemcpyAsync(dv.p, v, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemcpyAsync(df.p, from, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
	leaf_mask_kernel<<<1, 128, 0, c->stream>>>(T, logsrc, dm.p);
	ray_sees_batch_kernel<<<blocks_for(n, 128), 128, 0, c->stream>>>(T, dm.p, dp.p, dv.p, df.p, logsrc, n, dh.p);
	c->launches += 2;
	CUDA_TRY(cudaMemcpyAsync(sees, dh.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	CUDA_TRY(cudaGetLastError());
	return RADK_OK;
}

int radk_vertex_ff(radk_ctx *c, int shooter, const int *receiver, const int *vertex, int n, float *ff) {
	if (c && c->split_mode) { set_error("radk_vertex_ff: not available on adaptive-split contexts"); return RADK_ERR_STATE; }
	if (!c || !receiver || !vertex || !ff || n < 0) { set_error("radk_vertex_ff: bad argument"); return RADK_ERR_ARG; }
	if (shooter < 0 || shooter >= c->trinum || c->h_slot_to_recv[shooter] < 0) {
		set_error("radk_vertex_ff: shooter slot is not a receiver of this context");
		return RADK_ERR_ARG;
	}
	if (n == 0) return RADK_OK;
	CUDA_TRY(cudaSetDevice(c->device));
	std::vector<int> recv(n), vert(n), src(1, c->h_slot_to_recv[shooter]);
	for (int i = 0; i < n; i++) {
		if (vertex[i] < 0 || vertex[i] > 2 || receiver[i] < 0 || receiver[i] >= c->trinum) { set_error("radk_vertex_ff: bad (receiver, vertex) pair"); return RADK_ERR_ARG; }
		recv[i] = c->h_slot_to_recv[receiver[i]], vert[i] = vertex[i];
	}
	DevBuf<int> dr, dv, ds;
	DevBuf<float> df;
	CUDA_TRY(dr.upload(recv, c->stream));
	CUDA_TRY(dv.upload(vert, c->stream));
	CUDA_TRY(ds.upload(src, c->stream));
	CUDA_TRY(df.alloc(n));
	ShooterRec *tmp = c->rec.p + 2 + kMaxSources - 1;
	make_sources_kernel<<<1, 32, 0, c->stream>>>(c->S, ds.p, 1, tmp, 0);
	vertex_ff_batch_kernel<<<blocks_for(n, 128), 128, 0, c->stream>>>(c->S, tmp, dr.p, dv.p, n, df.p);
	c->launches += 2;
	CUDA_TRY(cudaMemcpyAsync(ff, df.p, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	CUDA_TRY(cudaGetLastError());
	return RADK_OK;
}

int radk_ctx_peer_handle(radk_ctx *c, void *handle) {
	if (c && c->split_mode) { set_error("radk_ctx_peer_handle: not available on adaptive-split contexts"); return RADK_ERR_STATE; }
	if (!c || !handle) { set_error("radk_ctx_peer_handle: null argument"); return RADK_ERR_ARG; }
	if (!c->use_fast) { set_error("radk_ctx_peer_handle: context does not use the batched path"); return RADK_ERR_STATE; }
	static_assert(sizeof(cudaIpcMemHandle_t) == RADK_PEER_HANDLE_BYTES, "handle size");
	CUDA_TRY(cudaSetDevice(c->device));
	cudaIpcMemHandle_t h;
	CUDA_TRY(cudaIpcGetMemHandle(&h, c->box.p));
	memcpy(handle, &h, sizeof(h));
	return RADK_OK;
}

int radk_ctx_peer_connect(radk_ctx *c, const void *handles, int world) {
	if (!c || !handles || world != c->opt.world) { set_error("radk_ctx_peer_connect: bad argument"); return RADK_ERR_ARG; }
	if (!c->use_fast) { set_error("radk_ctx_peer_connect: context does not use the batched path"); return RADK_ERR_STATE; }
	CUDA_TRY(cudaSetDevice(c->device));
	for (int t = 0; t < world; t++) {
		if (t == c->opt.rank) continue;
		cudaIpcMemHandle_t h;
		memcpy(&h, (const char *)handles + (size_t)t * RADK_PEER_HANDLE_BYTES, sizeof(h));
		void *p = nullptr;
		CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		c->ipc_opened.push_back(p);
		c->F.box[t] = (PeerBox *)p;
	}
	c->peers_connected = true;
	return RADK_OK;
}

int radk_ctx_peer_connect_local(radk_ctx *c, radk_ctx *const *peers, int world) {
	if (!c || !peers || world != c->opt.world) { set_error("radk_ctx_peer_connect_local: bad argument"); return RADK_ERR_ARG; }
	if (!c->use_fast) { set_error("radk_ctx_peer_connect_local: context does not use the batched path"); return RADK_ERR_STATE; }
	CUDA_TRY(cudaSetDevice(c->device));
	for (int t = 0; t < world; t++) {
		if (!peers[t] || peers[t]->opt.rank != t || peers[t]->opt.world != world || !peers[t]->use_fast) {
			set_error("radk_ctx_peer_connect_local: peers[r] must be rank r of the same world");
			return RADK_ERR_ARG;
		}
		if (t != c->opt.rank && peers[t]->device == c->device) {
			// ranks that share a device (tests): graphs with device-side loops of two contexts do not overlap on one
			// device, and each rank's kernels wait for the other's mail - launch the rounds from the host instead, and
			// on ONE stream per rank: with more streams than hardware queues a kernel that waits for a peer's mail can
			// end up queued in front of the peer's kernel that sends it
			c->graph_tried = true;
			c->overlap = false;
		}
		if (peers[t]->device != c->device) {
			cudaError_t e = cudaDeviceEnablePeerAccess(peers[t]->device, 0);
			if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
			else if (e != cudaSuccess) { set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return RADK_ERR_CUDA; }
		}
		c->F.box[t] = peers[t]->box.p;
	}
	c->peers_connected = true;
	return RADK_OK;
}

int radk_ctx_owned_slots(radk_ctx *c, int *slots, int cap, int *n) {
	if (c && c->split_mode) { set_error("radk_ctx_owned_slots: not available on adaptive-split contexts"); return RADK_ERR_STATE; }
	if (!c || !n) { set_error("radk_ctx_owned_slots: null argument"); return RADK_ERR_ARG; }
	*n = c->R;
	if (slots)
		for (int r = 0; r < c->R && r < cap; r++) slots[r] = c->recv_slot[r];
	return RADK_OK;
}

}  // extern "C"
