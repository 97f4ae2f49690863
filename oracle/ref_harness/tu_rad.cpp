// TEST INFRASTRUCTURE (oracle/_ref harness). Not part of the product.
//
// Translation unit 1 of 3: textually includes the UNMODIFIED reference driver
// (/root/reference/src/rad.cpp) so that its file-static step function
// `parallelRadiosity` (rad.cpp:406), `listCanditate` (rad.cpp:387) and the
// anonymous-namespace options (rad.cpp:24-31) are reachable, then drives
//   InitModelFromFile -> InitRadiosity -> PrePartitionTriangles -> N x parallelRadiosity
// exactly as `DoRadiosity` (rad.cpp:491-523) does, dumping raw state between.
//
// The reference path is given on the compiler command line (-DREF_RAD_CPP=...),
// nothing from the reference is copied into this repository.
#include <vector>
#include <string>
#include <stdint.h>

#define main ref_main
#include REF_RAD_CPP
#undef main

#include "harness_api.h"

static void dump_scene(const char *path, int n0)
{
	FILE *f = fopen(path, "wb");
	if (!f) { fprintf(stderr, "harness: cannot open %s\n", path); exit(2); }
	HarnessSceneHeader h;
	memset(&h, 0, sizeof(h));
	memcpy(h.magic, "RADSCN1", 8);
	h.n0 = n0;
	h.trinum = trinum;
	h.nnodes = harness_tree_nodes();
	h.nlist = harness_tree_listlen();
	h.gridnum = harness_grid_num();
	h.mingridlen = harness_min_grid_len();
	for (int i = 0; i < 3; i++) h.G0[i] = G0[i], h.G1[i] = G1[i];
	h.converge = ConvergeLimit;
	h.parallel_src = ParallelSource;
	h.light = LightScale;
	h.tri_limit = TriangleLimit;
	h.area_limit = AreaLimit;
	h.sample_area = SampleArea;
	h.delta_ff = DeltaFFLimit;
	h.tri_record_bytes = (int)sizeof(Triangle);
	h.node_record_bytes = (int)sizeof(TreeNode);
	fwrite(&h, sizeof(h), 1, f);
	fwrite(TriStore, sizeof(Triangle), trinum, f);
	fwrite(harness_tree_node_ptr(), sizeof(TreeNode), h.nnodes, f);
	fwrite(harness_tree_list_ptr(), sizeof(int), h.nlist, f);
	fclose(f);
}

int main(int argc, char *argv[])
{
	// harness-only options are "--xxx"; everything else is handed to the
	// reference's own ProcessOption (rad.cpp:525).
	std::vector<char*> ref_argv;
	ref_argv.push_back(argv[0]);
	const char *scene_path = NULL, *final_path = NULL, *shooters_path = NULL, *rays_path = NULL;
	long max_shots = -1;	// -1: to convergence
	long ray_log_shots = 0;
	int no_prepartition = 0, write_fin = 0, json = 0;
	for (int i = 1; i < argc; i++) {
		std::string a(argv[i]);
		if (a == "--dump-scene") scene_path = argv[++i];
		else if (a == "--dump-final") final_path = argv[++i];
		else if (a == "--dump-shooters") shooters_path = argv[++i];
		else if (a == "--log-rays") { rays_path = argv[++i]; ray_log_shots = atol(argv[++i]); }
		else if (a == "--shots") max_shots = atol(argv[++i]);
		else if (a == "--no-prepartition") no_prepartition = 1;
		else if (a == "--write-fin") write_fin = 1;
		else if (a == "--json") json = 1;
		else ref_argv.push_back(argv[i]);
	}
	omp_set_num_threads(ParallelJobs);
	FILE *fin = NULL;
	const char *prefix = ProcessOption((int)ref_argv.size(), ref_argv.data(), fin);
	if (fin == NULL) { fprintf(stderr, "harness: cannot open model\n"); return 1; }
	if (rays_path) omp_set_num_threads(1);	// the ray log is not thread-safe

	double t0 = omp_get_wtime();
	InitModelFromFile(fin);
	fclose(fin);
	int n0 = trinum;
	double t1 = omp_get_wtime();
	InitRadiosity();
	double t2 = omp_get_wtime();
	if (!no_prepartition) {
#ifdef PRE_PARTITION
		PrePartitionTriangles();
#endif
	}
	double t3 = omp_get_wtime();
	if (scene_path) dump_scene(scene_path, n0);

	FILE *fs = shooters_path ? fopen(shooters_path, "w") : NULL;
	if (rays_path) harness_ray_log_open(rays_path);
	harness_ray_counters_reset();
	long shots = 0;
	double sel_s = 0, loop_s = 0;
	double tl0 = omp_get_wtime();
	while (max_shots < 0 || shots < max_shots) {
		if (rays_path && shots >= ray_log_shots) harness_ray_log_close();
		if (fs) {
			double ts = omp_get_wtime();
			std::vector<int> C(listCanditate());
			sel_s += omp_get_wtime() - ts;
			for (size_t j = 0; j < C.size(); j++) {
				TrianglePtr tp = &TriStore[C[j]];
				float d = tp->deltaB[0] + tp->deltaB[1] + tp->deltaB[2];
				union { float f; uint32_t u; } cv; cv.f = d;
				fprintf(fs, "%ld %d %08x %d\n", shots, C[j], cv.u, tp->parent);
			}
		}
		if (parallelRadiosity() == 0)
			break;
		shots++;
	}
	loop_s = omp_get_wtime() - tl0 - sel_s * (fs ? 1 : 0);
	harness_ray_log_close();
	if (fs) fclose(fs);
	if (final_path) dump_scene(final_path, n0);
	if (write_fin) PrintOut(prefix, 0);

	long long rays = 0, visible = 0;
	harness_ray_counters(&rays, &visible);
	int leaves = 0;
	for (int i = 0; i < trinum; i++) leaves += TriStore[i].parent >= 0;
	if (json) {
		printf("HARNESS_JSON {\"n0\": %d, \"trinum\": %d, \"leaves\": %d, \"shots\": %ld, "
			"\"rays\": %lld, \"visible\": %lld, \"loop_s\": %.6f, \"read_build_s\": %.6f, "
			"\"init_s\": %.6f, \"prepartition_s\": %.6f, \"threads\": %d, \"nodes\": %d, \"listlen\": %d, "
			"\"gridnum\": %d, \"mingridlen\": %.9g}\n",
			n0, trinum, leaves, shots, rays, visible, loop_s, t1 - t0, t2 - t1, t3 - t2,
			rays_path ? 1 : ParallelJobs, harness_tree_nodes(), harness_tree_listlen(),
			harness_grid_num(), (double)harness_min_grid_len());
	}
	return 0;
}
