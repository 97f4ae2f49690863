// TEST / DESIGN INFRASTRUCTURE - warp-level cost simulation of the form-factor walk (not shipped).
//
// Replays the rays ff_batch_kernel traces for a few shooters of a model on the CPU (same rt_core.cuh code),
// records each ray's leaf visits (list length, tests executed, mode) and then replays them through two
// warp scheduling schemes with a simple issue-slot cost model:
//   seq : the shipped scheme - every lane scans its own leaf list sequentially (lanes wait for the longest list)
//   coop: cooperative scheme - lanes push their leaf's entries into a per-warp queue and all 32 lanes test
//         queue items (no lane waits on another lane's long list)
// Output: issue slots per ray and lane utilisation for each scheme.  Used to choose the kernel structure
// before GPU time is spent.
//
// build: g++ -O2 -std=c++17 -ffp-contract=off -I include -I hw-radiosity_b200/csrc -I hw-radiosity_b200/host
//            tests/hostemu/simt_sim.cpp -o /tmp/simt_sim -L hw-radiosity_b200 -lradk -Wl,-rpath,$PWD/hw-radiosity_b200
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "radk.h"
#include "rt_core.cuh"
#include "scene_flatten.hpp"

using namespace radk;

struct Visit {
	int n;		// list entries (excluding ban)
	int tests;	// tests the sequential scan executes (early exit in any-hit leaves)
	int full;	// of those, tests that go past the plane part (side planes evaluated)
	bool nearest;	// leaf lists the shooter's original
	bool hit;	// the walk ends here
};
struct Ray {
	std::vector<Visit> v;
};

static int g_chunk = 4, g_srep = 3, g_sthr = 20, g_tthr = 99;
static long long cat_slots[5][4], cat_useful[5][4];

int main(int argc, char **argv) {
	const char *path = argc > 1 ? argv[1] : "/tmp/church.tri";
	const int nshooters = argc > 2 ? atoi(argv[2]) : 6;
	g_chunk = argc > 3 ? atoi(argv[3]) : 4;
	radk_model_opts mo;
	radk_default_model_opts(&mo);
	radk_model *m = nullptr;
	if (radk_model_load(path, &mo, &m)) { fprintf(stderr, "load failed: %s\n", radk_last_error()); return 1; }
	radk_scene s;
	radk_model_scene(m, &s);
	FlatScene F;
	if (flatten_scene(&s, 0, 1, F)) { fprintf(stderr, "%s\n", F.error.c_str()); return 1; }
	const DevTree T = F.host_tree(&s);
	// shooters: the receivers with the largest unshot energy
	std::vector<std::pair<float, int>> cand;
	for (int r = 0; r < F.R; r++) cand.push_back({F.st[0][r].w + F.st[1][r].w + F.st[2][r].w, r});
	std::sort(cand.rbegin(), cand.rend());
	const int tile = argc > 4 ? atoi(argv[4]) : 1024;
	g_srep = argc > 5 ? atoi(argv[5]) : 3, g_sthr = argc > 6 ? atoi(argv[6]) : 20; g_tthr = argc > 7 ? atoi(argv[7]) : 99;
	long long nrays = 0, nvisits = 0, ntests = 0, nentries = 0,
		  nempty = 0, nnearest = 0;
	long long hist[9] = {0};
	// cost model (warp instructions)
	const int C_FETCH = 25, C_LOOKUP = 60, C_NEXT = 45, C_TEST_PLANE = 38, C_TEST_SIDE = 30, C_LIST_OVH = 6;
	const int C_EXPAND = 30, C_ITEM = 50, C_ITEM_SIDE = 30, C_RESOLVE = 12;
	for (int si = 0; si < nshooters; si++) {
		const int r = cand[si * 7 % cand.size()].second;	// spread over the candidate list
		const float p[3][3] = {{F.geo[0][r].x, F.geo[0][r].y, F.geo[0][r].z}, {F.geo[1][r].x, F.geo[1][r].y, F.geo[1][r].z},
				       {F.geo[2][r].x, F.geo[2][r].y, F.geo[2][r].z}};
		const int parent = F.recv_parent[r];
		const float4 pn = F.par_n[parent];
		ShooterRec sh;
		make_shooter(p, pn.x, pn.y, pn.z, pn.w, F.geo[2][r].w, 1, 1, 1, parent, F.recv_slot[r], r, &sh);
		std::vector<unsigned> mask((T.nnodes + 31) / 32, 0u);
		for (int k = T.tri_leaf_begin[sh.logsrc]; k < T.tri_leaf_begin[sh.logsrc + 1]; k++)
			mask[T.tri_leaves[k] >> 5] |= 1u << (T.tri_leaves[k] & 31);
		for (int h = 0; h < 2; h++) {
			const FFSample &smp = sh.half[h];
			for (int base = 0; base < F.U; base += tile) {
				// the CTA's ray queue
				std::vector<Ray> q;
				for (int u = base; u < std::min(F.U, base + tile); u++) {
					const float4 v = F.uv[u];
					const int ban = float_as_int_hd(v.w);
					const float4 n = F.par_n[ban];
					float dx, dy, dz;
					bool cast;
					ff_geom(smp, n.x, n.y, n.z, n.w, v.x, v.y, v.z, dx, dy, dz, cast);
					if (!cast) continue;
					Ray ray;
					const float px = v.x, py = v.y, pz = v.z, vx = dx, vy = dy, vz = dz;
					float t = 0, qx = px, qy = py, qz = pz;
					int prev = -1, cur, steps = 0, guard = 0;
					while ((cur = leaf_of<false>(T, qx, qy, qz, nullptr)) >= 0) {
						if (++steps > kMaxWalkSteps) break;
						if (prev == cur) {
							t = (float)((double)t + 1.0);
							qx = px + t * vx, qy = py + t * vy, qz = pz + t * vz;
							if (++guard > 100000) break;
							Visit vv = {0, 0, 0, false, false};
							ray.v.push_back(vv);
							continue;
						}
						const float4 lo = F.leaf_lo[cur], hi = F.leaf_hi[cur];
						Visit vv = {0, 0, 0, false, false};
						vv.nearest = (mask[cur >> 5] >> (cur & 31)) & 1u;
						const int4 *l4 = &F.list4[F.leaf_list4[cur]];
						bool done = false, anyhit = false;
						float t_min = 1000000.0f;
						for (; !done; l4++) {
							const int ids[4] = {l4->x, l4->y, l4->z, l4->w};
							for (int e = 0; e < 4; e++) {
								if (ids[e] < 0) { done = true; break; }
								const int idx = ids[e] & kIdMask;
								if (idx == ban) continue;
								vv.n++;
								if (anyhit && !vv.nearest) continue;	// sequential scan would have stopped
								vv.tests++;
								// plane part outcome
								const float4 *o = occ_plane(T, idx);
								const float a1 = dot3(vx, vy, vz, o[0].x, o[0].y, o[0].z);
								const float a2 = dot3(o[0].w - px, o[1].x - py, o[1].y - pz, o[0].x, o[0].y, o[0].z);
								bool plane_ok = !((a1 > 0.0f && a2 <= 0.0f) || (a1 < 0.0f && a2 >= 0.0f));
								if (plane_ok) { const float tt = a2 / a1; plane_ok = !(tt <= 0.0001f) && !(tt >= 1000000.0f); }
								if (plane_ok && !(ids[e] & kPlaneFlag)) vv.full++;
								float tm = vv.nearest ? t_min : 1000000.0f;
								if (tri_hit(o, occ_side(T, idx), px, py, pz, vx, vy, vz, tm)) { anyhit = true; if (vv.nearest) t_min = tm; }
							}
						}
						vv.hit = anyhit;
						ray.v.push_back(vv);
						if (anyhit) break;
						prev = cur;
						float tn = 1000000.0f, ss;
						if (fabsf(vx) > 1e-6f) { ss = ((vx > 0 ? hi.x : lo.x) - px) / vx; tn = (ss < tn) ? ss : tn; }
						if (fabsf(vy) > 1e-6f) { ss = ((vy > 0 ? hi.y : lo.y) - py) / vy; tn = (ss < tn) ? ss : tn; }
						if (fabsf(vz) > 1e-6f) { ss = ((vz > 0 ? hi.z : lo.z) - pz) / vz; tn = (ss < tn) ? ss : tn; }
						tn = (float)((double)tn + 0.1);
						qx = px + tn * vx, qy = py + tn * vy, qz = pz + tn * vz;
						t = tn;
					}
					nrays++;
					for (auto &vv : ray.v) {
						nvisits++, ntests += vv.tests, nentries += vv.n, nempty += vv.n == 0, nnearest += vv.nearest;
						hist[std::min(8, (vv.n + 3) / 4)]++;
					}
					q.push_back(std::move(ray));
				}
				// ---- replay: 8 warps x 32 lanes pull rays from the queue
				for (int scheme = 0; scheme < 5; scheme++) {
					size_t next = 0;
					struct Lane { int ray = -1; size_t vi = 0; int consumed = 0; bool pending = false; };
					std::vector<Lane> L(256);
					bool alive[8];
					for (int w = 0; w < 8; w++) alive[w] = true;
					bool any = true;
					long long *cs = cat_slots[scheme], *cu = cat_useful[scheme];
					while (any) {
						any = false;
						for (int w = 0; w < 8; w++) {
							if (!alive[w]) continue;
							Lane *l = &L[w * 32];
							if (scheme == 0) {
								// one outer iteration: fetch, lookup, list scan (max over lanes), nextpoint
								int nf = 0, na = 0, maxc = 0, nn = 0;
								long long sumc = 0;
								for (int k = 0; k < 32; k++) {
									if (l[k].ray < 0) { if (next < q.size()) l[k].ray = (int)next++, l[k].vi = 0, nf++; else { nf++; continue; } }
									na++;
									const Visit &vv = q[l[k].ray].v[l[k].vi];
									const int c = vv.tests * (C_TEST_PLANE + C_LIST_OVH) + vv.full * C_TEST_SIDE;
									maxc = std::max(maxc, c), sumc += c;
									if (!vv.hit) nn++;
									l[k].vi++;
									if (l[k].vi >= q[l[k].ray].v.size()) l[k].ray = -1;
								}
								if (na == 0) { alive[w] = false; continue; }
								any = true;
								cs[0] += (nf ? C_FETCH : 0), cu[0] += nf * C_FETCH;
								cs[1] += C_LOOKUP, cu[1] += na * C_LOOKUP;
								cs[2] += maxc, cu[2] += sumc;
								cs[3] += (nn ? C_NEXT : 0), cu[3] += nn * C_NEXT;
							} else if (scheme == 4) {
								// seq with repeated stepping: lanes whose leaf is empty step again (up to g_srep S phases) while at
								// least g_sthr lanes want to; then every lane scans its list sequentially
								int na = 0;
								for (int rep = 0; rep < g_srep; rep++) {
									int ns = 0, nf = 0, nn = 0;
									for (int k = 0; k < 32; k++) {
										if (l[k].pending) continue;
										if (l[k].ray < 0) { if (next < q.size()) l[k].ray = (int)next++, l[k].vi = 0, nf++; else continue; }
										ns++;
										if (l[k].vi > 0) nn++;
										const Visit &vv = q[l[k].ray].v[l[k].vi];
										if (vv.tests > 0) l[k].pending = true;
										else {
											l[k].vi++;
											if (l[k].vi >= q[l[k].ray].v.size()) l[k].ray = -1;
										}
									}
									if (ns == 0) break;
									cs[0] += (nf ? C_FETCH : 0), cu[0] += nf * C_FETCH;
									cs[1] += C_LOOKUP, cu[1] += ns * C_LOOKUP;
									cs[3] += (nn ? C_NEXT : 0), cu[3] += nn * C_NEXT;
									int nwant = 0;
									for (int k = 0; k < 32; k++) nwant += !l[k].pending && (l[k].ray >= 0 || next < q.size());
									if (nwant < g_sthr) break;
								}
								int maxc = 0;
								long long sumc = 0;
								for (int k = 0; k < 32; k++) {
									if (!l[k].pending) continue;
									na++;
									const Visit &vv = q[l[k].ray].v[l[k].vi];
									const int c = vv.tests * (C_TEST_PLANE + C_LIST_OVH) + vv.full * C_TEST_SIDE;
									maxc = std::max(maxc, c), sumc += c;
									l[k].pending = false;
									l[k].vi++;
									if (l[k].vi >= q[l[k].ray].v.size()) l[k].ray = -1;
								}
								bool act = false;
								for (int k = 0; k < 32; k++) act = act || l[k].ray >= 0;
								if (na == 0 && !act && next >= q.size()) { alive[w] = false; continue; }
								any = true;
								cs[2] += maxc, cu[2] += sumc;
							} else if (scheme == 3) {
								// dynamic: S when >= g_sthr lanes need it (or nobody tests), else one test step for every lane
								// that has entries left (early exit kept per lane)
								int nS = 0, nT = 0, nidle = 0;
								for (int k = 0; k < 32; k++) {
									if (l[k].pending) nT++;
									else if (l[k].ray >= 0 || next < q.size()) nS++;
									else nidle++;
								}
								if (nS == 0 && nT == 0) { alive[w] = false; continue; }
								any = true;
								if (nS >= g_sthr || nT == 0) {
									int nf = 0, ns = 0;
									for (int k = 0; k < 32; k++) {
										if (l[k].pending) continue;
										if (l[k].ray < 0) { if (next < q.size()) l[k].ray = (int)next++, l[k].vi = 0, nf++; else continue; }
										ns++;
										const Visit &vv = q[l[k].ray].v[l[k].vi];
										l[k].consumed = 0;
										if (vv.tests > 0) l[k].pending = true;
										else {
											l[k].vi++;
											if (l[k].vi >= q[l[k].ray].v.size()) l[k].ray = -1;
										}
									}
									cs[0] += (nf ? C_FETCH : 0), cu[0] += nf * C_FETCH;
									cs[1] += C_LOOKUP + C_NEXT + 8, cu[1] += ns * (C_LOOKUP + C_NEXT + 8);
								} else {
									int nside = 0, nt = 0;
									for (int k = 0; k < 32; k++) {
										if (!l[k].pending) continue;
										const Visit &vv = q[l[k].ray].v[l[k].vi];
										nt++;
										// side planes for the last `full` tests of the visit (approximation)
										if (l[k].consumed >= vv.tests - vv.full) nside++;
										l[k].consumed++;
										if (l[k].consumed >= vv.tests) {
											l[k].pending = false;
											l[k].vi++;
											if (l[k].vi >= q[l[k].ray].v.size()) l[k].ray = -1;
										}
									}
									const int c = C_TEST_PLANE + C_LIST_OVH + 6;
									cs[2] += c + (nside ? C_TEST_SIDE : 0), cu[2] += nt * c + nside * C_TEST_SIDE;
								}
							} else {
								// S phase for lanes without a pending chunk (scheme 2: repeated while many lanes sit in empty
								// leaves); then expansion + T over the warp queue
								int npend = 0;
								for (int rep = 0; rep < (scheme == 2 ? g_srep : 1); rep++) {
									int ns = 0, nf = 0;
									for (int k = 0; k < 32; k++) {
										if (!l[k].pending) {
											if (l[k].ray < 0) { nf++; if (next < q.size()) l[k].ray = (int)next++, l[k].vi = 0; else continue; }
											ns++;
											const Visit &vv = q[l[k].ray].v[l[k].vi];
											l[k].consumed = 0;
											if (vv.n > 0) l[k].pending = true;
											else {
												l[k].vi++;
												if (l[k].vi >= q[l[k].ray].v.size()) l[k].ray = -1;
											}
										}
									}
									cs[0] += (nf ? C_FETCH : 0), cu[0] += nf * C_FETCH;
									cs[1] += ns ? (C_LOOKUP + C_NEXT) : 0, cu[1] += ns * (C_LOOKUP + C_NEXT);
									npend = 0;
									for (int k = 0; k < 32; k++) npend += l[k].pending;
									if (npend >= g_sthr) break;
								}
								bool first = true;
								for (;;) {
								int items = 0, sides = 0;
								npend = 0;
								for (int k = 0; k < 32; k++) npend += l[k].pending;
								if (!first && npend < g_tthr) break;
								for (int k = 0; k < 32; k++) {
									if (!l[k].pending) continue;
									const Visit &vv = q[l[k].ray].v[l[k].vi];
									const int take = std::min(g_chunk, vv.n - l[k].consumed);
									items += take;
									sides += (int)((long long)vv.full * take / std::max(1, vv.tests));
									l[k].consumed += take;
									const bool found = vv.hit && !vv.nearest && l[k].consumed >= vv.tests;
									if (l[k].consumed >= vv.n || found) {
										l[k].pending = false;
										l[k].vi++;
										if (l[k].vi >= q[l[k].ray].v.size()) l[k].ray = -1;
									}
								}
								if (first) {
									bool act = false;
									for (int k = 0; k < 32; k++) act = act || l[k].ray >= 0 || l[k].pending;
									if (npend == 0 && !act) { alive[w] = false; break; }
									any = true;
								}
								first = false;
								const int iters = (items + 31) / 32;
								cs[3] += npend ? C_EXPAND + C_RESOLVE : 0, cu[3] += npend * (C_EXPAND + C_RESOLVE);
								cs[2] += iters * C_ITEM + ((sides + 31) / 32) * C_ITEM_SIDE, cu[2] += items * C_ITEM + sides * C_ITEM_SIDE;
								if (g_tthr > 32) break;
								}
							}
						}
					}
				}
			}
		}
	}
	printf("rays %lld  visits/ray %.2f  entries/visit %.2f  tests/ray %.2f  empty visits %.1f%%  nearest visits %.1f%%\n", nrays,
	       (double)nvisits / nrays, (double)nentries / nvisits, (double)ntests / nrays, 100.0 * nempty / nvisits, 100.0 * nnearest / nvisits);
	printf("list length histogram (chunks of 4: 0,1..7,8+):");
	for (int k = 0; k < 9; k++) printf(" %.1f%%", 100.0 * hist[k] / nvisits);
	printf("\n");
	const char *names[5] = {"seq ", "coop", "coopS", "dyn ", "seqS"};
	const char *cats[4] = {"fetch", "lookup(+next)", "tests", "next|expand"};
	for (int sc = 0; sc < 5; sc++) {
		long long ts = 0, tu = 0;
		for (int k = 0; k < 4; k++) ts += cat_slots[sc][k], tu += cat_useful[sc][k];
		printf("%s: %.1f issue slots/ray, lane utilisation %.2f |", names[sc], (double)ts / nrays, (double)tu / 32 / ts);
		for (int k = 0; k < 4; k++) printf(" %s %.1f (%.2f)", cats[k], (double)cat_slots[sc][k] / nrays, cat_slots[sc][k] ? (double)cat_useful[sc][k] / 32 / cat_slots[sc][k] : 0.0);
		printf("\n");
	}
	radk_model_free(m);
	return 0;
}
