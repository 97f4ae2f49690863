// DESIGN INFRASTRUCTURE (not shipped): statistics of the visibility walk that ff_batch_kernel runs, gathered on the CPU
// with the same rt_core.cuh code.  For a few shooters of a model it counts, per traced ray,
//   tests      triangle tests of the bounded any-hit scan as shipped,
//   distinct   tests left when an occluder is tested at most once per ray (exact: a rejected occluder stays rejected,
//              see DESIGN.md "walk"),
//   notprev    tests left when only the occluders of the immediately preceding leaf are skipped,
// and the leaf-visit histogram.  Used to decide on list layouts before GPU time is spent.
//
// build: g++ -O2 -std=c++17 -ffp-contract=off -I include -I hw-radiosity_b200/csrc -I hw-radiosity_b200/host
//            tools/walk_stats.cpp -o /tmp/walk_stats -L hw-radiosity_b200 -lradk -Wl,-rpath,$PWD/hw-radiosity_b200
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <vector>

#include "radk.h"
#include "rt_core.cuh"
#include "scene_flatten.hpp"

using namespace radk;

int main(int argc, char **argv) {
	const char *path = argc > 1 ? argv[1] : "/tmp/models/church.tri";
	const int nshooters = argc > 2 ? atoi(argv[2]) : 6;
	radk_model_opts mo;
	radk_default_model_opts(&mo);
	radk_model *m = nullptr;
	if (radk_model_load(path, &mo, &m)) { fprintf(stderr, "load failed: %s\n", radk_last_error()); return 1; }
	radk_scene s;
	radk_model_scene(m, &s);
	FlatScene F;
	if (flatten_scene(&s, 0, 1, F)) { fprintf(stderr, "%s\n", F.error.c_str()); return 1; }
	const DevTree T = F.host_tree(&s);
	std::vector<std::pair<float, int>> cand;
	for (int r = 0; r < F.R; r++) cand.push_back({F.st[0][r].w + F.st[1][r].w + F.st[2][r].w, r});
	std::sort(cand.rbegin(), cand.rend());
	long long nrays = 0, nvisits = 0, ntests = 0, ndistinct = 0, nnotprev = 0, nempty = 0, nvis = 0, nblocked_first = 0;
	long long plane_reject_sign = 0, plane_reject_t = 0, side_evals = 0, accepts = 0;
	long long visit_hist[16] = {0};
	long long pair_same_path = 0, pair_total = 0;
	for (int si = 0; si < nshooters; si++) {
		const int r = cand[si * 7 % cand.size()].second;
		const float p[3][3] = {{F.geo[0][r].x, F.geo[0][r].y, F.geo[0][r].z}, {F.geo[1][r].x, F.geo[1][r].y, F.geo[1][r].z},
				       {F.geo[2][r].x, F.geo[2][r].y, F.geo[2][r].z}};
		const int parent = F.recv_parent[r];
		const float4 pn = F.par_n[parent];
		ShooterRec sh;
		make_shooter(p, pn.x, pn.y, pn.z, pn.w, F.geo[2][r].w, 1, 1, 1, parent, F.recv_slot[r], r, &sh);
		std::vector<unsigned> mask((T.nnodes + 31) / 32, 0u);
		for (int k = T.tri_leaf_begin[sh.logsrc]; k < T.tri_leaf_begin[sh.logsrc + 1]; k++)
			mask[T.tri_leaves[k] >> 5] |= 1u << (T.tri_leaves[k] & 31);
		std::vector<std::vector<int>> paths[2];
		for (int h = 0; h < 2; h++) {
			const FFSample &smp = sh.splittable ? sh.half[h] : sh.centre;
			paths[h].assign(F.U, {});
			for (int u = 0; u < F.U; u++) {
				const float4 v = F.uv[u];
				const int ban = float_as_int_hd(v.w);
				const float4 n = F.par_n[ban];
				float dx, dy, dz;
				bool cast;
				ff_geom(smp, n.x, n.y, n.z, n.w, v.x, v.y, v.z, dx, dy, dz, cast);
				if (!cast) continue;
				const float px = v.x, py = v.y, pz = v.z, vx = dx, vy = dy, vz = dz;
				float t = 0, qx = px, qy = py, qz = pz;
				int prev = -1, cur, steps = 0, guard = 0, visits = 0;
				std::set<int> seen, prevset;
				bool visible = false;
				std::vector<int> &path = paths[h][u];
				path.push_back(-2);
				while ((cur = leaf_of<false>(T, qx, qy, qz, nullptr)) >= 0) {
					if (++steps > kMaxWalkSteps) break;
					if (prev == cur) {
						t = (float)((double)t + 1.0);
						qx = px + t * vx, qy = py + t * vy, qz = pz + t * vz;
						if (++guard > 100000) break;
						continue;
					}
					path.push_back(cur);
					visits++;
					const float4 lo = F.leaf_lo[cur], hi = F.leaf_hi[cur];
					const bool lists_src = (mask[cur >> 5] >> (cur & 31)) & 1u;
					// the shipped scan: logsrc alone first, then bounded any-hit over the list, stop at first blocker
					float t_max = 1000000.0f;
					int tie = -1;
					bool src_hit = false;
					if (lists_src && sh.logsrc != ban) {
						ntests++;
						if (!seen.count(sh.logsrc)) ndistinct++;
						if (!prevset.count(sh.logsrc)) nnotprev++;
						src_hit = tri_hit(occ_plane(T, sh.logsrc), occ_side(T, sh.logsrc), px, py, pz, vx, vy, vz, t_max);
						if (src_hit) tie = sh.logsrc;
					}
					std::set<int> thisset;
					const int4 *l4 = &F.list4[F.leaf_list4[cur]];
					bool done = false, blocked = false;
					int nent = 0;
					for (; !done; l4++) {
						const int ids[4] = {l4->x, l4->y, l4->z, l4->w};
						for (int e = 0; e < 4; e++) {
							if (ids[e] < 0) { done = true; break; }
							const int idx = ids[e] & kIdMask;
							thisset.insert(idx);
							if (idx == ban) continue;
							nent++;
							if (blocked) continue;
							ntests++;
							if (!seen.count(idx)) ndistinct++;
							if (!prevset.count(idx)) nnotprev++;
							seen.insert(idx);
							const float4 *o = occ_plane(T, idx);
							const float a1 = dot3(vx, vy, vz, o[0].x, o[0].y, o[0].z);
							const float a2 = dot3(o[0].w - px, o[1].x - py, o[1].y - pz, o[0].x, o[0].y, o[0].z);
							if ((a1 > 0.0f && a2 <= 0.0f) || (a1 < 0.0f && a2 >= 0.0f)) plane_reject_sign++;
							else {
								const float tt = a2 / a1;
								if (tt <= 0.0001f || tt > t_max) plane_reject_t++;
								else if (!(ids[e] & kPlaneFlag)) side_evals++;
							}
							const int rr = any_hit_bounded_one(T, ids[e], px, py, pz, vx, vy, vz, ban, t_max, tie);
							if (rr) blocked = true, accepts++;
						}
					}
					if (nent == 0) nempty++;
					prevset = thisset;
					if (blocked) { if (visits == 1) nblocked_first++; break; }
					if (src_hit) { visible = true; break; }
					prev = cur;
					float tn = 1000000.0f, ss;
					if (fabsf(vx) > 1e-6f) { ss = ((vx > 0 ? hi.x : lo.x) - px) / vx; tn = (ss < tn) ? ss : tn; }
					if (fabsf(vy) > 1e-6f) { ss = ((vy > 0 ? hi.y : lo.y) - py) / vy; tn = (ss < tn) ? ss : tn; }
					if (fabsf(vz) > 1e-6f) { ss = ((vz > 0 ? hi.z : lo.z) - pz) / vz; tn = (ss < tn) ? ss : tn; }
					tn = (float)((double)tn + 0.1);
					qx = px + tn * vx, qy = py + tn * vy, qz = pz + tn * vz;
					t = tn;
				}
				nrays++, nvisits += visits, nvis += visible;
				visit_hist[std::min(15, visits)]++;
			}
		}
		for (int u = 0; u < F.U; u++) {
			if (paths[0][u].empty() || paths[1][u].empty()) continue;
			pair_total++;
			if (paths[0][u] == paths[1][u]) pair_same_path++;
		}
	}
	printf("%s: %d shooters, rays %lld (visible %.2f%%, blocked in first leaf %.1f%%)\n", path, nshooters, nrays, 100.0 * nvis / nrays,
	       100.0 * nblocked_first / nrays);
	printf("  per ray: visits %.2f (empty %.1f%%), tests %.2f, distinct %.2f, not-in-previous-leaf %.2f\n", (double)nvisits / nrays,
	       100.0 * nempty / nvisits, (double)ntests / nrays, (double)ndistinct / nrays, (double)nnotprev / nrays);
	printf("  tests: sign reject %.1f%%, t reject %.1f%%, side planes evaluated %.1f%%, accepted %.1f%%\n", 100.0 * plane_reject_sign / ntests,
	       100.0 * plane_reject_t / ntests, 100.0 * side_evals / ntests, 100.0 * accepts / ntests);
	printf("  vertices with both halves cast: %lld, identical leaf sequence: %.1f%%\n", pair_total, 100.0 * pair_same_path / std::max(1ll, pair_total));
	printf("  visits histogram:");
	for (int i = 0; i < 16; i++) printf(" %d:%.1f%%", i, 100.0 * visit_hist[i] / nrays);
	printf("\n");
	return 0;
}
