// TEST INFRASTRUCTURE - CPU emulation of the device code paths.
//
// Compiles hw-radiosity_b200/csrc/rt_core.cuh (the __host__ __device__ core: octree walk, triangle
// test, form factor, shooter setup) and scene_flatten.hpp with plain g++ so that the logic the
// kernels run can be checked against the oracle in the CPU-only test tier, before any GPU time is
// spent.  It re-states the few lines of kernel glue (per-thread bodies of vertex_ff_kernel and
// update_select_kernel) sequentially.  It is NOT part of the product and NOT a fallback: nothing
// under hw-radiosity_b200/ references this file.
#include <cstdio>
#include <cstring>
#include <vector>

#include "radk.h"
#include "rt_core.cuh"
#include "scene_flatten.hpp"

using namespace radk;

namespace {
inline unsigned ordered_bits(float f) {
	unsigned u; memcpy(&u, &f, 4);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
}

extern "C" {

int emu_rays(const radk_scene *s, const float *p, const float *v, const int *from, int n, int *out) {
	FlatScene F;
	if (flatten_scene(s, 0, 1, F)) { fprintf(stderr, "emu: %s\n", F.error.c_str()); return 1; }
	const DevTree T = F.host_tree(s);
	for (int i = 0; i < n; i++)
		out[i] = ray_walk<false>(T, p[3 * i], p[3 * i + 1], p[3 * i + 2], v[3 * i], v[3 * i + 1], v[3 * i + 2], from[i], nullptr);
	return 0;
}

// ray_sees (bounded any-hit leaf visits, the walk of ff_batch_kernel) for n rays, each toward its own shooter
// original logsrc[i]: out[i] = 1 iff the walk ends on logsrc[i].  nan_visits counts rays whose walk accepted a NaN t
// (the order-dependent case that falls back to the sequential scan).
int emu_sees(const radk_scene *s, const float *p, const float *v, const int *from, const int *logsrc, int n, int *out) {
	FlatScene F;
	if (flatten_scene(s, 0, 1, F)) { fprintf(stderr, "emu: %s\n", F.error.c_str()); return 1; }
	const DevTree T = F.host_tree(s);
	std::vector<unsigned> mask((T.nnodes + 31) / 32);
	for (int i = 0; i < n; i++) {
		std::fill(mask.begin(), mask.end(), 0u);
		for (int k = T.tri_leaf_begin[logsrc[i]]; k < T.tri_leaf_begin[logsrc[i] + 1]; k++)
			mask[T.tri_leaves[k] >> 5] |= 1u << (T.tri_leaves[k] & 31);
		out[i] = ray_sees(T, mask.data(), p[3 * i], p[3 * i + 1], p[3 * i + 2], v[3 * i], v[3 * i + 1], v[3 * i + 2], from[i], logsrc[i]);
	}
	return 0;
}

// Runs `shots` single-source shots the way the kernels do (unique-vertex form factors, then the
// per-receiver update with the fused argmax) and writes accB[trinum][3][3], deltaB[trinum][3] and
// the shooter slot of every shot.  counters[0..2] = reference-equivalent rays, visible, shaded receivers.
static int g_mode = 0;	// 0: cal_ff / ray_walk (per-shot kernels); 1: ff_geom + ray_sees (leaf_visit, the walk of ff_batch_kernel)
void emu_set_mode(int mode) { g_mode = mode; }

int emu_run(const radk_scene *s, float converge, int rank, int world, int shots, float *accB, float *deltaB,
	    int *shooters, int *nshots_out, long long *counters) {
	FlatScene F;
	if (flatten_scene(s, rank, world, F)) { fprintf(stderr, "emu: %s\n", F.error.c_str()); return 1; }
	const DevTree T = F.host_tree(s);
	const int R = F.R, U = F.U;
	std::vector<float> ff0(U), ff1(U);
	std::vector<unsigned> vstat(U);
	auto select = [&](ShooterRec *out) {
		unsigned long long best = 0;
		for (int r = 0; r < R; r++) {
			const float delta = F.st[0][r].w + F.st[1][r].w + F.st[2][r].w;
			if (!(delta > converge)) continue;
			const unsigned long long k = ((unsigned long long)ordered_bits(delta) << 32) | (unsigned)F.recv_slot[r];
			best = k > best ? k : best;
		}
		if (!best) { out->valid = 0; return; }
		const int r = F.slot_to_recv[(int)(best & 0xffffffffu)];
		const float p[3][3] = {{F.geo[0][r].x, F.geo[0][r].y, F.geo[0][r].z}, {F.geo[1][r].x, F.geo[1][r].y, F.geo[1][r].z},
				       {F.geo[2][r].x, F.geo[2][r].y, F.geo[2][r].z}};
		const int parent = F.recv_parent[r];
		const float4 pn = F.par_n[parent];
		make_shooter(p, pn.x, pn.y, pn.z, pn.w, F.geo[2][r].w, F.st[0][r].w, F.st[1][r].w, F.st[2][r].w, parent,
			     F.recv_slot[r], r, out);
	};
	long long rays = 0, visible = 0, shaded = 0;
	int done = 0;
	ShooterRec sh;
	select(&sh);
	std::vector<unsigned> mask((T.nnodes + 31) / 32);
	while (done < shots && sh.valid) {
		if (g_mode == 1) {	// ff_batch_kernel: per-row leaf mask, culls first, then the any-hit walk
			std::fill(mask.begin(), mask.end(), 0u);
			for (int k = T.tri_leaf_begin[sh.logsrc]; k < T.tri_leaf_begin[sh.logsrc + 1]; k++)
				mask[T.tri_leaves[k] >> 5] |= 1u << (T.tri_leaves[k] & 31);
			for (int u = 0; u < U; u++) {
				const float4 v = F.uv[u];
				const int parent = float_as_int_hd(v.w);
				const float4 pn = F.par_n[parent];
				unsigned nr = 0, nv = 0;
				float out[2] = {0.0f, 0.0f};
				const int nh = sh.splittable ? 2 : 1;
				for (int h = 0; h < nh; h++) {
					const FFSample &smp = sh.splittable ? sh.half[h] : sh.centre;
					float dx, dy, dz;
					bool cast;
					const float ff = ff_geom(smp, pn.x, pn.y, pn.z, pn.w, v.x, v.y, v.z, dx, dy, dz, cast);
					if (!cast) continue;
					nr++;
					if (ray_sees(T, mask.data(), v.x, v.y, v.z, dx, dy, dz, parent, sh.logsrc)) out[h] = ff, nv++;
				}
				if (sh.splittable) {	// the discarded centre sample, for reference-equivalent counts only
					float dx, dy, dz;
					bool cast;
					ff_geom(sh.centre, pn.x, pn.y, pn.z, pn.w, v.x, v.y, v.z, dx, dy, dz, cast);
					if (cast) nr++, nv += ray_sees(T, mask.data(), v.x, v.y, v.z, dx, dy, dz, parent, sh.logsrc);
				}
				ff0[u] = out[0], ff1[u] = out[1];
				vstat[u] = nr | (nv << 8);
			}
		} else
		for (int u = 0; u < U; u++) {	// vertex_ff_kernel
			const float4 v = F.uv[u];
			const int parent = float_as_int_hd(v.w);
			const float4 pn = F.par_n[parent];
			FFCounters fc = {0, 0, 0};
			WalkCounters wc = {0, 0, 0, 0, 0, 0, 0};
			if (sh.splittable) {
				// reference-equivalent counting incl. the discarded centre sample
				cal_ff<true>(T, sh.centre, sh.logsrc, pn.x, pn.y, pn.z, pn.w, parent, v.x, v.y, v.z, &fc, &wc);
				ff0[u] = cal_ff<true>(T, sh.half[0], sh.logsrc, pn.x, pn.y, pn.z, pn.w, parent, v.x, v.y, v.z, &fc, &wc);
				ff1[u] = cal_ff<true>(T, sh.half[1], sh.logsrc, pn.x, pn.y, pn.z, pn.w, parent, v.x, v.y, v.z, &fc, &wc);
			} else {
				ff0[u] = cal_ff<true>(T, sh.centre, sh.logsrc, pn.x, pn.y, pn.z, pn.w, parent, v.x, v.y, v.z, &fc, &wc);
				ff1[u] = 0.0f;
			}
			vstat[u] = fc.rays | (fc.visible << 8);
		}
		for (int r = 0; r < R; r++) {	// update_select_kernel
			float4 &a0 = F.st[0][r], &a1 = F.st[1][r], &a2 = F.st[2][r];
			if (F.recv_slot[r] == sh.slot) { a0.w = a1.w = a2.w = 0.0f; continue; }
			if (!(a0.x < 255.0f || a0.y < 255.0f || a0.z < 255.0f || a1.x < 255.0f || a1.y < 255.0f || a1.z < 255.0f ||
			      a2.x < 255.0f || a2.y < 255.0f || a2.z < 255.0f)) continue;
			shaded++;
			const int4 id = F.vid[r];
			const int ids[3] = {id.x, id.y, id.z};
			float ff[3];
			for (int v = 0; v < 3; v++) {
				ff[v] = ff0[ids[v]] + ff1[ids[v]];
				rays += vstat[ids[v]] & 0xff, visible += vstat[ids[v]] >> 8;
			}
			const float4 fk = F.par_fk[id.w];
			const float lo[3] = {fk.x * sh.dB[0] / 3, fk.y * sh.dB[1] / 3, fk.z * sh.dB[2] / 3};
			float dB[3] = {a0.w, a1.w, a2.w};
			float4 *acc[3] = {&a0, &a1, &a2};
			for (int v = 0; v < 3; v++) {
				float d;
				d = lo[0] * ff[v]; dB[0] += d; acc[v]->x += d;
				d = lo[1] * ff[v]; dB[1] += d; acc[v]->y += d;
				d = lo[2] * ff[v]; dB[2] += d; acc[v]->z += d;
			}
			a0.w = dB[0], a1.w = dB[1], a2.w = dB[2];
		}
		if (shooters) shooters[done] = sh.slot;
		done++;
		select(&sh);
	}
	for (int r = 0; r < R; r++) {
		const int slot = F.recv_slot[r];
		for (int v = 0; v < 3; v++) {
			accB[slot * 9 + v * 3 + 0] = F.st[v][r].x;
			accB[slot * 9 + v * 3 + 1] = F.st[v][r].y;
			accB[slot * 9 + v * 3 + 2] = F.st[v][r].z;
			deltaB[slot * 3 + v] = F.st[v][r].w;
		}
	}
	if (nshots_out) *nshots_out = done;
	if (counters) counters[0] = rays, counters[1] = visible, counters[2] = shaded, counters[3] = U, counters[4] = R;
	return 0;
}

}  // extern "C"

// ---- experiment: how many cast rays start in a leaf with a guaranteed plane blocker? ----------------
extern "C" int emu_start_block_stats(const radk_scene *s, int shooter_slot, long long *out) {
	FlatScene F;
	if (flatten_scene(s, 0, 1, F)) return 1;
	const DevTree T = F.host_tree(s);
	const int r = F.slot_to_recv[shooter_slot];
	const float p[3][3] = {{F.geo[0][r].x, F.geo[0][r].y, F.geo[0][r].z}, {F.geo[1][r].x, F.geo[1][r].y, F.geo[1][r].z},
			       {F.geo[2][r].x, F.geo[2][r].y, F.geo[2][r].z}};
	const int parent = F.recv_parent[r];
	const float4 pn = F.par_n[parent];
	ShooterRec sh;
	make_shooter(p, pn.x, pn.y, pn.z, pn.w, F.geo[2][r].w, 1, 1, 1, parent, shooter_slot, r, &sh);
	std::vector<unsigned> mask((T.nnodes + 31) / 32, 0u);
	for (int k = T.tri_leaf_begin[sh.logsrc]; k < T.tri_leaf_begin[sh.logsrc + 1]; k++)
		mask[T.tri_leaves[k] >> 5] |= 1u << (T.tri_leaves[k] & 31);
	const float ext = s->G1[0] - s->G0[0];
	const double eps = 1e-4 * ext;
	long long cast = 0, killed = 0, blocked = 0, first_leaf_blocked = 0;
	for (int h = 0; h < 2; h++) {
		const FFSample &smp = sh.half[h];
		std::vector<char> sb(T.nnodes, 0);
		for (int leaf = 0; leaf < T.nnodes; leaf++) {
			if (F.leaf_list4[leaf] < 0 || ((mask[leaf >> 5] >> (leaf & 31)) & 1u)) continue;
			const float4 lo = F.leaf_lo[leaf], hi = F.leaf_hi[leaf];
			for (const int4 *l4 = &F.list4[F.leaf_list4[leaf]];; l4++) {
				const int ids[4] = {l4->x, l4->y, l4->z, l4->w};
				bool end = false;
				for (int e = 0; e < 4 && !end; e++) {
					if (ids[e] < 0) { end = true; break; }
					if (!(ids[e] & kPlaneFlag)) continue;
					const float4 *o = &F.occ[(size_t)(ids[e] & kIdMask) * OCC_PLANE];
					const double n[3] = {o[0].x, o[0].y, o[0].z}, c[3] = {o[0].w, o[1].x, o[1].y};
					const double blo[3] = {lo.x, lo.y, lo.z}, bhi[3] = {hi.x, hi.y, hi.z};
					double pmin = 0, pmax = 0;	// range of p.n over the cell box
					for (int k = 0; k < 3; k++) {
						pmin += n[k] > 0 ? n[k] * blo[k] : n[k] * bhi[k];
						pmax += n[k] > 0 ? n[k] * bhi[k] : n[k] * blo[k];
					}
					const double d = c[0] * n[0] + c[1] * n[1] + c[2] * n[2];
					const double sdot = smp.cx * n[0] + smp.cy * n[1] + smp.cz * n[2];
					// both the plane and the sample strictly above, or both strictly below, the whole box
					double hp_min, hp_max, hs_min, hs_max;
					if (d > pmax && sdot > pmax) hp_min = d - pmax, hp_max = d - pmin, hs_min = sdot - pmax, hs_max = sdot - pmin;
					else if (d < pmin && sdot < pmin) hp_min = pmin - d, hp_max = pmax - d, hs_min = pmin - sdot, hs_max = pmax - sdot;
					else continue;
					if (hp_min < eps || hs_min < eps) continue;
					if (hp_min < 2.5e-4 * (hs_max + eps)) continue;
					if (hp_max > 1e5 * hs_min) continue;
					sb[leaf] = 1;
				}
				if (end || sb[leaf]) break;
			}
		}
		for (int u = 0; u < F.U; u++) {
			const float4 v = F.uv[u];
			const int par = float_as_int_hd(v.w);
			const float4 n4 = F.par_n[par];
			float dx, dy, dz;
			bool c;
			ff_geom(smp, n4.x, n4.y, n4.z, n4.w, v.x, v.y, v.z, dx, dy, dz, c);
			if (!c) continue;
			cast++;
			const int leaf = leaf_of<false>(T, v.x, v.y, v.z, nullptr);
			const bool sees = ray_sees(T, mask.data(), v.x, v.y, v.z, dx, dy, dz, par, sh.logsrc);
			if (!sees) blocked++;
			if (leaf >= 0 && sb[leaf]) {
				killed++;
				if (sees) return 2;	// the criterion must never kill a visible ray
			}
			if (leaf >= 0 && !((mask[leaf >> 5] >> (leaf & 31)) & 1u) && !sees) {
				if (leaf_visit(T, leaf, false, v.x, v.y, v.z, dx, dy, dz, par, sh.logsrc) == 0) first_leaf_blocked++;
			}
		}
	}
	out[0] = cast, out[1] = blocked, out[2] = first_leaf_blocked, out[3] = killed;
	return 0;
}

// ---- stateful shard emulation of the chained shot loop (fast_kernels.cuh) ------------------------------
// One EmuShard = one rank: its receiver shard, its form-factor cache, the (replicated) cache directory.  A round is
//   emu_shard_top       local top candidates (key + complete shooter record)            [predict_kernel]
//   emu_shard_begin     adopt the global top list: rebuild the cache, fill the new rows  [predict_merge + ff_batch]
//   emu_shard_cands     records of the candidates this rank owns                         [chain_kernel, first half]
//   emu_shard_chain     schedule from the complete candidate records                     [chain_kernel]
//   emu_shard_apply     apply pass, returns the rank's first violation                   [apply_kernel<false>]
//   emu_shard_commit    global first violation -> valid prefix, re-apply if needed       [apply_kernel<true>]
// The exchanges between the calls are done by the caller: gloo all_gathers in tests/test_multi_rank_cpu.py, plain
// loops in emu_run_chain below.  The candidate records, the schedule and the per-receiver passes run the code of
// shot_core.cuh that the kernels run.
#include "shot_core.cuh"
#include <algorithm>

struct EmuShard {
	FlatScene F;
	DevTree T;
	int rows = 0;
	float converge = 1200.0f;
	std::vector<int> row_slot;		// cache directory (identical on every rank)
	std::vector<ShooterRec> rec;		// shooter record per row
	std::vector<float> ff;			// [rows][U][2]
	std::vector<unsigned> mask;
	std::vector<float4> prev[3];
	ShotSched sched[kMaxChain];
	int nsched = 0;
	int parallel_src = 1;			// -parallel_src: co-shooters per iteration
	long long local_count = 0;		// receivers of this shard above the convergence limit (emu_shard_top)
	long long rows_filled = 0, rounds = 0, reapplied = 0;
};

extern "C" void *emu_shard_open(const radk_scene *s, int rank, int world, int rows, float converge) {
	EmuShard *e = new EmuShard();
	if (flatten_scene(s, rank, world, e->F)) { delete e; return nullptr; }
	e->T = e->F.host_tree(s);
	e->rows = rows, e->converge = converge;
	e->row_slot.assign(rows, -1);
	e->rec.resize(rows);
	e->ff.assign((size_t)rows * e->F.U * 2, 0.0f);
	e->mask.resize((e->T.nnodes + 31) / 32);
	for (int v = 0; v < 3; v++) e->prev[v].resize(e->F.R);
	return e;
}
extern "C" void emu_shard_close(void *h) { delete (EmuShard *)h; }
extern "C" int emu_shard_sizes(void *h, int *R, int *U) { EmuShard *e = (EmuShard *)h; *R = e->F.R, *U = e->F.U; return 0; }
extern "C" void emu_shard_set_parallel_src(void *h, int P) { ((EmuShard *)h)->parallel_src = P < 1 ? 1 : P; }
extern "C" long long emu_shard_count(void *h) { return ((EmuShard *)h)->local_count; }
// div3 (shot_core.cuh) against the division it replaces: n patterns from `start` with stride `step`; returns mismatches
extern "C" long long emu_div3_check(unsigned start, unsigned step, long long n) {
	long long bad = 0;
	unsigned u = start;
	for (long long i = 0; i < n; i++, u += step) {
		float x;
		memcpy(&x, &u, 4);
		if (x != x) continue;
		volatile float xv = x;
		const float a = xv / 3.0f, b = div3(x);
		if (memcmp(&a, &b, 4) != 0) bad++;
	}
	return bad;
}
extern "C" int emu_shooter_rec_bytes(void) { return (int)sizeof(ShooterRec); }
extern "C" int emu_cand_rec_bytes(void) { return (int)sizeof(CandRec); }

static void emu_make_rec(const FlatScene &F, int r, ShooterRec *rec) {
	const float p[3][3] = {{F.geo[0][r].x, F.geo[0][r].y, F.geo[0][r].z}, {F.geo[1][r].x, F.geo[1][r].y, F.geo[1][r].z},
			       {F.geo[2][r].x, F.geo[2][r].y, F.geo[2][r].z}};
	const float4 pn = F.par_n[F.recv_parent[r]];
	make_shooter(p, pn.x, pn.y, pn.z, pn.w, F.geo[2][r].w, F.st[0][r].w, F.st[1][r].w, F.st[2][r].w, F.recv_parent[r],
		     F.recv_slot[r], r, rec);
}

// local top `rows` candidates, descending; keys[i] = 0 past the end
extern "C" int emu_shard_top(void *h, unsigned long long *keys, ShooterRec *recs) {
	EmuShard *e = (EmuShard *)h;
	const FlatScene &F = e->F;
	std::vector<unsigned long long> all;
	for (int r = 0; r < F.R; r++) {
		const unsigned long long k = energy_key(e->converge, F.st[0][r].w, F.st[1][r].w, F.st[2][r].w, F.recv_slot[r]);
		if (k) all.push_back(k);
	}
	e->local_count = (long long)all.size();
	const size_t n = std::min<size_t>(all.size(), (size_t)e->rows);
	std::partial_sort(all.begin(), all.begin() + n, all.end(), std::greater<unsigned long long>());
	for (int i = 0; i < e->rows; i++) {
		keys[i] = (size_t)i < n ? all[i] : 0ull;
		memset(&recs[i], 0, sizeof(ShooterRec));
		if (keys[i]) emu_make_rec(F, F.slot_to_recv[(int)(keys[i] & 0xffffffffu)], &recs[i]);
	}
	return 0;
}

// global top list (descending keys, 0 = none) with the shooter records: rebuild the cache like rebuild_cache() does and
// fill the new rows with this shard's form factors.  Returns the number of rows filled.
extern "C" int emu_shard_begin(void *h, const unsigned long long *keys, const ShooterRec *recs) {
	EmuShard *e = (EmuShard *)h;
	FlatScene &F = e->F;
	const DevTree &T = e->T;
	const int rows = e->rows;
	std::vector<int> cand_row(rows, -1);
	std::vector<char> keep(rows, 0);
	for (int i = 0; i < rows; i++) {
		if (!keys[i]) continue;
		const int slot = (int)(keys[i] & 0xffffffffu);
		for (int j = 0; j < rows; j++)
			if (e->row_slot[j] == slot) cand_row[i] = j, keep[j] = 1;
	}
	for (int j = 0; j < rows; j++)
		if (!keep[j]) e->row_slot[j] = -1;
	std::vector<int> todo;
	int next_free = 0;
	for (int i = 0; i < rows; i++) {
		if (!keys[i] || cand_row[i] >= 0) continue;
		while (next_free < rows && e->row_slot[next_free] >= 0) next_free++;
		if (next_free >= rows) break;
		e->row_slot[next_free] = (int)(keys[i] & 0xffffffffu);
		e->rec[next_free] = recs[i];
		todo.push_back(next_free);
	}
	for (int row : todo) {
		const ShooterRec &sh = e->rec[row];
		std::fill(e->mask.begin(), e->mask.end(), 0u);
		for (int k = T.tri_leaf_begin[sh.logsrc]; k < T.tri_leaf_begin[sh.logsrc + 1]; k++)
			e->mask[T.tri_leaves[k] >> 5] |= 1u << (T.tri_leaves[k] & 31);
		float *out = &e->ff[(size_t)row * F.U * 2];
		for (int u = 0; u < F.U; u++) {
			const float4 v = F.uv[u];
			const int parent = float_as_int_hd(v.w);
			const float4 pn = F.par_n[parent];
			out[2 * u] = out[2 * u + 1] = 0.0f;
			const int nh = sh.splittable ? 2 : 1;
			for (int hh = 0; hh < nh; hh++) {
				const FFSample &smp = sh.splittable ? sh.half[hh] : sh.centre;
				float dx, dy, dz;
				bool cast;
				const float ff = ff_geom(smp, pn.x, pn.y, pn.z, pn.w, v.x, v.y, v.z, dx, dy, dz, cast);
				if (cast && ray_sees(T, e->mask.data(), v.x, v.y, v.z, dx, dy, dz, parent, sh.logsrc)) out[2 * u + hh] = ff;
			}
		}
	}
	e->rows_filled += (long long)todo.size(), e->rounds++;
	return (int)todo.size();
}

struct EmuFF {
	const float *ff;
	int U;
	float operator()(int row, int u) const { return ff[((size_t)row * U + u) * 2] + ff[((size_t)row * U + u) * 2 + 1]; }
};

// records of the candidates (rows) this rank owns; owned[i] = 0 for rows that are free or owned elsewhere
extern "C" int emu_shard_cands(void *h, CandRec *out, int *owned) {
	EmuShard *e = (EmuShard *)h;
	const FlatScene &F = e->F;
	const EmuFF ffa{e->ff.data(), F.U};
	for (int i = 0; i < e->rows; i++) {
		memset(&out[i], 0, sizeof(CandRec));
		out[i].slot = -1, owned[i] = 0;
		const int slot = e->row_slot[i];
		const int r = slot >= 0 ? F.slot_to_recv[slot] : -1;
		if (r < 0) continue;
		owned[i] = 1;
		CandRec &c = out[i];
		c.a0 = F.st[0][r], c.a1 = F.st[1][r], c.a2 = F.st[2][r];
		const int4 id = F.vid[r];
		const float4 fk = F.par_fk[id.w];
		c.fk[0] = fk.x, c.fk[1] = fk.y, c.fk[2] = fk.z, c.slot = slot;
		for (int j = 0; j < e->rows; j++) {
			const bool filled = e->row_slot[j] >= 0;
			c.fsum[j][0] = filled ? ffa(j, id.x) : 0.0f, c.fsum[j][1] = filled ? ffa(j, id.y) : 0.0f, c.fsum[j][2] = filled ? ffa(j, id.z) : 0.0f;
		}
	}
	return 0;
}

// count = receivers above the convergence limit over all ranks (sum of emu_shard_count), budget in iterations
extern "C" int emu_shard_chain2(void *h, const CandRec *cands, long long budget, long long count) {
	EmuShard *e = (EmuShard *)h;
	std::vector<CandRec> c(cands, cands + e->rows);
	e->nsched = chain_schedule_host(c.data(), e->rows, e->converge, budget, e->parallel_src, count, e->sched);
	return e->nsched;
}
extern "C" int emu_shard_chain(void *h, const CandRec *cands, long long budget) { return emu_shard_chain2(h, cands, budget, 0); }

extern "C" int emu_shard_apply(void *h) {
	EmuShard *e = (EmuShard *)h;
	FlatScene &F = e->F;
	const EmuFF ffa{e->ff.data(), F.U};
	int viol = e->nsched;
	for (int r = 0; r < F.R; r++) {
		for (int v = 0; v < 3; v++) e->prev[v][r] = F.st[v][r];
		const int4 id = F.vid[r];
		const float4 fk = F.par_fk[id.w];
		const int v1 = apply_schedule(F.st[0][r], F.st[1][r], F.st[2][r], id.x, id.y, id.z, fk.x, fk.y, fk.z, F.recv_slot[r], e->sched,
					      e->nsched, e->converge, ffa, e->parallel_src > 1);
		viol = std::min(viol, v1);
	}
	return viol;
}

// viol = minimum over the ranks; returns the number of committed iterations and the first shooter of each (what the
// reference's loop counter / the shot log see)
extern "C" int emu_shard_commit(void *h, int viol, int *slots) {
	EmuShard *e = (EmuShard *)h;
	FlatScene &F = e->F;
	const int K = e->nsched, n = viol < K ? viol + 1 : K;
	if (n < K) {
		const EmuFF ffa{e->ff.data(), F.U};
		for (int r = 0; r < F.R; r++) {
			for (int v = 0; v < 3; v++) F.st[v][r] = e->prev[v][r];
			const int4 id = F.vid[r];
			const float4 fk = F.par_fk[id.w];
			apply_schedule(F.st[0][r], F.st[1][r], F.st[2][r], id.x, id.y, id.z, fk.x, fk.y, fk.z, F.recv_slot[r], e->sched, n, e->converge, ffa, e->parallel_src > 1);
		}
		e->reapplied++;
	}
	int iters = 0;
	for (int k = 0; k < n; k++)
		if (e->sched[k].flags & kSchedHead) slots[iters++] = e->sched[k].slot;
	return iters;
}

// owned slots and their state: slots[R], accB[R][9], deltaB[R][3]
extern "C" int emu_shard_state(void *h, int *slots, float *accB, float *deltaB) {
	EmuShard *e = (EmuShard *)h;
	FlatScene &F = e->F;
	for (int r = 0; r < F.R; r++) {
		slots[r] = F.recv_slot[r];
		for (int v = 0; v < 3; v++) {
			accB[r * 9 + v * 3 + 0] = F.st[v][r].x, accB[r * 9 + v * 3 + 1] = F.st[v][r].y, accB[r * 9 + v * 3 + 2] = F.st[v][r].z;
			deltaB[r * 3 + v] = F.st[v][r].w;
		}
	}
	return 0;
}

// The whole chained loop for `world` ranks in one process: accB[trinum][3][3], deltaB[trinum][3] by slot, shooter slots.
// counters: [0] rows filled (rank 0), [1] rounds, [2] rounds that re-applied a prefix.
static int g_parallel_src = 1;
extern "C" void emu_set_parallel_src(int P) { g_parallel_src = P < 1 ? 1 : P; }
extern "C" int emu_run_chain(const radk_scene *s, float converge, int world, int rows, int shots, float *accB, float *deltaB,
			      int *shooters, int *nshots_out, long long *counters) {
	std::vector<EmuShard *> sh(world);
	for (int w = 0; w < world; w++) {
		if (!(sh[w] = (EmuShard *)emu_shard_open(s, w, world, rows, converge))) return 1;
		emu_shard_set_parallel_src(sh[w], g_parallel_src);
	}
	int done = 0;
	std::vector<unsigned long long> keys((size_t)world * rows), top(rows);
	std::vector<ShooterRec> recs((size_t)world * rows), toprec(rows);
	std::vector<CandRec> cands(rows), mine(rows);
	std::vector<int> owned(rows), slots(kMaxChain);
	while (done < shots) {
		for (int w = 0; w < world; w++) emu_shard_top(sh[w], &keys[(size_t)w * rows], &recs[(size_t)w * rows]);
		std::vector<int> order((size_t)world * rows);
		for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
		std::sort(order.begin(), order.end(), [&](int a, int b) { return keys[a] > keys[b]; });
		for (int i = 0; i < rows; i++) top[i] = keys[order[i]], toprec[i] = recs[order[i]];
		if (!top[0]) break;	// converged
		for (int w = 0; w < world; w++) emu_shard_begin(sh[w], top.data(), toprec.data());
		for (int i = 0; i < rows; i++) memset(&cands[i], 0, sizeof(CandRec)), cands[i].slot = -1;
		for (int w = 0; w < world; w++) {
			emu_shard_cands(sh[w], mine.data(), owned.data());
			for (int i = 0; i < rows; i++)
				if (owned[i]) cands[i] = mine[i];
		}
		int K = 0, viol = kMaxChain + 1;
		long long count = 0;
		for (int w = 0; w < world; w++) count += sh[w]->local_count;
		for (int w = 0; w < world; w++) K = emu_shard_chain2(sh[w], cands.data(), shots - done, count);
		for (int w = 0; w < world; w++) viol = std::min(viol, emu_shard_apply(sh[w]));
		int n = 0;
		for (int w = 0; w < world; w++) n = emu_shard_commit(sh[w], viol, slots.data());
		if (K <= 0 || n <= 0) { for (auto e : sh) delete e; return 2; }	// the first scheduled shot is always valid
		for (int k = 0; k < n; k++) shooters[done + k] = slots[k];
		done += n;
	}
	for (int w = 0; w < world; w++) {
		const FlatScene &F = sh[w]->F;
		for (int r = 0; r < F.R; r++) {
			const int slot = F.recv_slot[r];
			for (int v = 0; v < 3; v++) {
				accB[slot * 9 + v * 3 + 0] = F.st[v][r].x, accB[slot * 9 + v * 3 + 1] = F.st[v][r].y, accB[slot * 9 + v * 3 + 2] = F.st[v][r].z;
				deltaB[slot * 3 + v] = F.st[v][r].w;
			}
		}
	}
	if (nshots_out) *nshots_out = done;
	if (counters) counters[0] = sh[0]->rows_filled, counters[1] = sh[0]->rounds, counters[2] = sh[0]->reapplied;
	for (auto e : sh) delete e;
	return 0;
}
