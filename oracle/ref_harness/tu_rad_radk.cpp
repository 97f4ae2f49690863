// TEST INFRASTRUCTURE (oracle/_ref harness). Not part of the product.
//
// INTEGRATION.md section 2, compiled: the UNMODIFIED reference driver (/root/reference/src/rad.cpp, included
// textually) keeps its own option parser, .tri reader, BuildTree, InitRadiosity, PrePartitionTriangles and PrintOut;
// only the loop of DoRadiosity (rad.cpp:497-516: `parallelRadiosity()` per iteration) is replaced by the
// radk_ctx_step calls the integration guide shows, linked against the product's libradk.so.  The octree accessor
// the guide adds to raycast.cpp is tu_raycast.cpp's (that file includes raycast.cpp unmodified as well).
// tests/test_reference_host.py compares every file this binary writes with what the reference's own binary
// (oracle/_ref/rad_oracle) writes.
#include <vector>
#include <string>
#include <stdint.h>

#define main ref_main
#include REF_RAD_CPP
#undef main

#include "harness_api.h"
#include "radk.h"

static void ExportTree(radk_scene *s)		// the file-statics of raycast.cpp:9-16
{
	s->nodes = (const radk_tree_node *)harness_tree_node_ptr();
	s->nnodes = harness_tree_nodes();
	s->list = harness_tree_list_ptr();
	s->nlist = harness_tree_listlen();
	s->min_grid_len = harness_min_grid_len();
	s->grid_num = harness_grid_num();
}

static std::vector<float> g_acc, g_dB;
static std::vector<int> g_par;
static void PullRadiosity(radk_ctx *ctx)	// device state -> the reference's TriStore, for its own PrintOut
{
	g_acc.resize((size_t)trinum * 9), g_dB.resize((size_t)trinum * 3), g_par.resize(trinum);
	if (radk_ctx_download(ctx, g_acc.data(), g_dB.data(), g_par.data(), trinum)) {
		fprintf(stderr, "%s\n", radk_last_error());
		exit(1);
	}
	for (int i = 0; i < trinum; i++)
		if (g_par[i] >= 0) {
			memcpy(TriStore[i].accB, &g_acc[(size_t)i * 9], 36);
			memcpy(TriStore[i].deltaB, &g_dB[(size_t)i * 3], 12);
		}
}

static void DoRadiosityOnDevice(const char fileName[])	// DoRadiosity (rad.cpp:491-523) with the patched loop
{
	static_assert(sizeof(Triangle) == sizeof(radk_triangle), "radk_triangle must mirror the reference's Triangle");
	static_assert(sizeof(TreeNode) == sizeof(radk_tree_node), "radk_tree_node must mirror the reference's TreeNode");
	int n0 = trinum;	// input triangles = octree occluders (BuildTree ran on them)
	InitRadiosity();
#ifdef PRE_PARTITION
	PrePartitionTriangles();
#endif
	radk_scene sc;
	radk_opts op;
	radk_ctx *ctx = NULL;
	memset(&sc, 0, sizeof(sc));
	sc.tri = (const radk_triangle *)TriStore, sc.n0 = n0, sc.trinum = trinum;
	for (int k = 0; k < 3; k++) sc.G0[k] = G0[k], sc.G1[k] = G1[k];
	ExportTree(&sc);
	radk_default_opts(&op);
	op.converge = ConvergeLimit, op.parallel_src = ParallelSource, op.tri_limit = TriangleLimit;
	if (radk_ctx_create(&sc, &op, &ctx)) { fprintf(stderr, "%s\n", radk_last_error()); exit(1); }
	long long shots = 0, done = 0;
	int converged = 0;
	while (!converged) {
		if (radk_ctx_step(ctx, WriteIteration - shots % WriteIteration, &done, &converged)) {
			fprintf(stderr, "%s\n", radk_last_error());
			exit(1);
		}
		shots += done;
		if (!converged && shots % WriteIteration == 0) {	// PrintOut(fileName, loop) cadence, rad.cpp:510-513
			PullRadiosity(ctx);
			PrintOut(fileName, (int)shots);
			fprintf(stderr, "[" KGRN "INFO" KWHT "] END Loop. There are %i triangles.\n", trinum);
		}
	}
	fprintf(stderr, "[" KGRN "INFO" KWHT "] Complete in #Iterations %lld\n", shots + 1);
	PullRadiosity(ctx);
	radk_ctx_destroy(ctx);
	PrintOut(fileName, 0);
}

int main(int argc, char *argv[])	// the reference's main (rad.cpp:635-661) with DoRadiosity swapped
{
	omp_set_num_threads(ParallelJobs);
	FILE *fin;
	const char *foutName = ProcessOption(argc, argv, fin);
	if (fin == NULL) {
		fprintf(stderr, "File open error::%s\n", argv[2]);
		exit(1);
	}
	InitModelFromFile(fin);
	fclose(fin);
	DoRadiosityOnDevice(foutName);
	return 0;
}
