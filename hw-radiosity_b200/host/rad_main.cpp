// `rad` - command-line driver with the reference's surface (src/rad.cpp:525-661): same options,
// same .tri input, same "<prefix>.<loop>" / "<prefix>.fin" outputs, same log lines.  The loop itself
// runs on the GPU through the C ABI of include/radk.h; this file contains no numerical code.
//
// Options accepted exactly as by the reference's ProcessOption (rad.cpp:527-528):
//   -debug -adapt_area -sample_area -converge -delta_ff -write_cycle -zip -triangle -light -o
//   -interactive -jobs -parallel_src
// Additions (not in the reference; defaults leave its behaviour unchanged):
//   -gpu <n>        CUDA device ordinal (default 0)
//   -max_tri <n>    raise the reference's compile-time MaxTri pool limit (rad.h:9)
//   -stats          print device ray statistics at the end
//   -gpus <n>       shard the receivers over n GPUs of this box (devices gpu .. gpu+n-1), one host thread each
//   -split          adaptive splitting inside the loop on the device (the reference's split-enabled build:
//                   config.h without DONT_PARTITION/PRE_PARTITION); -adapt_area/-sample_area/-delta_ff then act
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "radk.h"

#define KRED "\x1B[31m"
#define KGRN "\x1B[32m"
#define KMAG "\x1B[35m"
#define KWHT "\x1B[37m"

namespace {

struct Options {
	int debug = 0;
	float adapt_area = 5.0f, sample_area = -1.0f, converge = 1200.0f, delta_ff = 0.000005f;
	int triangle = 1000000;
	float light = 1.0f;
	int write_cycle = 10, zip = 0, interactive = 0, jobs = 2, parallel_src = 1;
	int gpu = 0, gpus = 1, max_tri = 0, stats = 0, split = 0;
	const char *input = nullptr;
	const char *prefix = "./test";
};

double now() {
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void usage(const char *argv0, const Options &d) {
	fprintf(stderr, "Usage: %s [options] input_file -o output_file\n", argv0);
	fprintf(stderr, "OPTIONS\n");
	fprintf(stderr, "   Debug Options\n");
	fprintf(stderr, "       -debug <integer>        Output intermediate result according debug level.\n");
	fprintf(stderr, "                               " KMAG "Default -debug %d\n\n" KWHT, d.debug);
	fprintf(stderr, "   Radiosity Options\n");
	fprintf(stderr, "       -adapt_area <float>     The threahold of adaptive splitting algorithm.\n");
	fprintf(stderr, "                               " KMAG "Default -area %f\n" KWHT, d.adapt_area);
	fprintf(stderr, "       -sample_area <float>    When the difference of form factor for each vertex is\n");
	fprintf(stderr, "                               greater than delta form factor, it should try to split.\n");
	fprintf(stderr, "                               " KMAG "Default by model-dependent\n" KWHT);
	fprintf(stderr, "       -converge <float>       The radiosity B is the energy per unit area unit B is too small.\n");
	fprintf(stderr, "                               " KMAG "Default -converge %f\n" KWHT, d.converge);
	fprintf(stderr, "       -delta_ff <float>       The difference of delta form factor which is smaller than delta_ff\n");
	fprintf(stderr, "                               will consider as the same.\n");
	fprintf(stderr, "                               " KMAG "Default -delta_ff %f\n" KWHT, d.delta_ff);
	fprintf(stderr, "       -write_cycle <integer>  Write the status of radiosity for each write_cycle iterations.\n");
	fprintf(stderr, "                               " KMAG "Default -write_cycle %d\n" KWHT, d.write_cycle);
	fprintf(stderr, "       -triangle <integer>     The maximum the number of triangles in the model. If you show image\n");
	fprintf(stderr, "                               on WebGL, set -triangle 30000 is the best resolution\n");
	fprintf(stderr, "                               " KMAG "Defulat -triangle %d\n" KWHT, d.triangle);
	fprintf(stderr, "       -light <float>          Adjust the scale of bright light for testing.\n");
	fprintf(stderr, "                               " KMAG "Defulat -light %f\n" KWHT, d.light);
	fprintf(stderr, "       -jobs <integer>         Multithreading OpenMP omp_set_num_threads\n");
	fprintf(stderr, "                               " KMAG "Defulat -jobs %d\n" KWHT, d.jobs);
	fprintf(stderr, "       -parallel_src <integer> The number of light which illuminate other triangles each step\n");
	fprintf(stderr, "                               " KMAG "Defulat -parallel_src %d\n" KWHT, d.parallel_src);
	fprintf(stderr, "   Output Options\n");
	fprintf(stderr, "       -zip                    Compress output file by zip.\n");
	fprintf(stderr, "                               " KMAG "Defulat false\n" KWHT);
	fprintf(stderr, "       -o </<path>/file>       Assign the prefix filename the output file\n");
	fprintf(stderr, "                               " KMAG "Defulat ./test\n" KWHT);
	fprintf(stderr, "       -interactive            Transfer the result into WebGL each write_cycle\n");
	fprintf(stderr, "                               " KMAG "Defulat false\n" KWHT);
	fprintf(stderr, "   Device Options (this implementation)\n");
	fprintf(stderr, "       -gpu <integer>          CUDA device ordinal.  -max_tri <integer>  triangle pool size.  -stats\n");
	fprintf(stderr, "       -gpus <integer>         Shard the receivers over this many GPUs.\n");
	fprintf(stderr, "       -split                  Adaptive splitting inside the loop (reference built without DONT_PARTITION).\n");
}

bool want_int(int argc, char **argv, int &i, int &dst, const char *what) {
	if (i + 1 >= argc || sscanf(argv[++i], "%d", &dst) != 1) {	// the reference asserts here
		fprintf(stderr, "%s Error\n", what);
		return false;
	}
	return true;
}
bool want_float(int argc, char **argv, int &i, float &dst, const char *what) {
	if (i + 1 >= argc || sscanf(argv[++i], "%f", &dst) != 1) {
		fprintf(stderr, "%s Error\n", what);
		return false;
	}
	return true;
}

// rad.cpp:525-631
bool process_options(int argc, char **argv, Options &o) {
	if (argc < 3) {
		usage(argv[0], o);
		return false;
	}
	for (int i = 1; i < argc; i++) {
		const std::string a(argv[i]);
		bool ok = true;
		if (a == "-debug") ok = want_int(argc, argv, i, o.debug, "Debug");
		else if (a == "-adapt_area") ok = want_float(argc, argv, i, o.adapt_area, "AreaLimit");
		else if (a == "-sample_area") ok = want_float(argc, argv, i, o.sample_area, "SampleArea");
		else if (a == "-converge") ok = want_float(argc, argv, i, o.converge, "ConvergeLimit");
		else if (a == "-delta_ff") ok = want_float(argc, argv, i, o.delta_ff, "DeltaFFLimit");
		else if (a == "-write_cycle") ok = want_int(argc, argv, i, o.write_cycle, "Debug WriteIteration");
		else if (a == "-zip") o.zip = 1;
		else if (a == "-triangle") ok = want_int(argc, argv, i, o.triangle, "TriangleLimit");
		else if (a == "-light") ok = want_float(argc, argv, i, o.light, "LightScale");
		else if (a == "-o") { if (i + 1 < argc) o.prefix = argv[++i]; }
		else if (a == "-interactive") o.interactive = 1;
		else if (a == "-jobs") ok = want_int(argc, argv, i, o.jobs, "Parallel Jobs");
		else if (a == "-parallel_src") ok = want_int(argc, argv, i, o.parallel_src, "Parallel Source");
		else if (a == "-gpu") ok = want_int(argc, argv, i, o.gpu, "GPU");
		else if (a == "-gpus") ok = want_int(argc, argv, i, o.gpus, "GPUs");
		else if (a == "-max_tri") ok = want_int(argc, argv, i, o.max_tri, "MaxTri");
		else if (a == "-stats") o.stats = 1;
		else if (a == "-split") o.split = 1;
		else o.input = argv[i];		// rad.cpp:625-626: anything unrecognised is the input file
		if (!ok) return false;
	}
	fprintf(stderr, "[" KGRN "INFO" KWHT "] Input Model " KMAG "%s" KWHT " and Output Prefix " KMAG "%s\n" KWHT,
		o.input ? o.input : "(null)", o.prefix);
	return true;
}

// PrintOut's side effects after the file is written (rad.cpp:224-241)
void after_write(const Options &o, const std::string &fn) {
	if (o.zip) {
		std::string cmd = "gzip " + fn;
		printf("Compress output file %s .\n", fn.c_str());
		if (system(cmd.c_str())) puts("OK");	// sic: the reference prints OK on a non-zero status
		else puts("Failed");
	}
	if (o.interactive) {
		std::string cmd = "cd ./WebGL_demo && ./run.sh ../" + fn + " 2>/dev/null";
		int ret = system(cmd.c_str());
		fprintf(stderr, "WebGL: %s, return %d\n", cmd.c_str(), ret);
	}
}

bool write_snapshot(const std::vector<radk_ctx *> &ctxs, radk_model *model, const Options &o, int loop) {
	radk_ctx *ctx = ctxs[0];
	if (o.split) {		// the pool grew on the device: fetch whole records
		int n = 0;
		radk_ctx_download_triangles(ctx, nullptr, 0, &n);
		std::vector<radk_triangle> all((size_t)n);
		if (radk_ctx_download_triangles(ctx, all.data(), n, &n) != RADK_OK || radk_model_set_triangles(model, all.data(), n) != RADK_OK) {
			fprintf(stderr, "[" KRED "ERR" KWHT "] %s\n", radk_last_error());
			return false;
		}
		if (radk_model_write(model, o.prefix, loop) != RADK_OK) return loop != 0;
		after_write(o, std::string(o.prefix) + (loop == 0 ? std::string(".fin") : "." + std::to_string(loop)));
		return true;
	}
	// frozen mesh: colours are reduced on the device (K5), the geometry text is cached, the file is written by a
	// background thread.  -zip / -interactive act on the finished file, so they wait for it.
	const int trinum = radk_ctx_trinum(ctx);
	static std::vector<int> rgb;
	rgb.resize((size_t)trinum * 3);
	if (radk_ctx_snapshot_colors(ctx, rgb.data(), trinum) != RADK_OK) {
		fprintf(stderr, "[" KRED "ERR" KWHT "] %s\n", radk_last_error());
		return false;
	}
	if (ctxs.size() > 1) {		// every rank reports its own receivers (-1 elsewhere): merge the shards
		static std::vector<int> part;
		part.resize(rgb.size());
		for (size_t r = 1; r < ctxs.size(); r++) {
			if (radk_ctx_snapshot_colors(ctxs[r], part.data(), trinum) != RADK_OK) {
				fprintf(stderr, "[" KRED "ERR" KWHT "] %s\n", radk_last_error());
				return false;
			}
			for (size_t k = 0; k < rgb.size(); k++)
				if (part[k] >= 0) rgb[k] = part[k];
		}
	}
	radk_model_write_colors(model, o.prefix, loop, rgb.data(), trinum);
	if (o.zip || o.interactive || loop == 0) {
		if (radk_model_write_wait(model) != RADK_OK) return loop != 0;
		after_write(o, std::string(o.prefix) + (loop == 0 ? std::string(".fin") : "." + std::to_string(loop)));
	}
	return true;
}

}  // namespace

int main(int argc, char **argv) {
	Options o;
	if (!process_options(argc, argv, o)) return 1;
	FILE *probe = o.input ? fopen(o.input, "r") : nullptr;
	if (!probe) {
		fprintf(stderr, "File open error::%s\n", argc > 2 ? argv[2] : "");	// rad.cpp:641-644
		return 1;
	}
	fclose(probe);

	radk_model_opts mo;
	radk_default_model_opts(&mo);
	mo.light_scale = o.light, mo.tri_limit = o.triangle, mo.sample_area = o.sample_area;
	if (o.split) mo.prepartition = 0;	// the split-enabled reference build has no PrePartition pass
	if (o.max_tri > 0) mo.max_tri = o.max_tri;
	radk_model *model = nullptr;
	double t0 = now();
	int rc = radk_model_load(o.input, &mo, &model);
	double t1 = now();
	if (rc != RADK_OK) {
		fprintf(stderr, "[" KRED "ERR" KWHT "] %s\n", radk_last_error());
		return 1;
	}
	radk_scene sc;
	radk_model_scene(model, &sc);
	fprintf(stderr, "[" KGRN "INFO" KWHT "] Read #Triangle " KMAG "%d" KWHT " from input file\n", sc.n0);
	fprintf(stderr, "[" KGRN "INFO" KWHT "] 3D-DDA #TreeNode = %d, #Trilist = %d\n", sc.nnodes, sc.nlist);
	fprintf(stderr, "[" KGRN "INFO" KWHT "] Init took " KMAG "%f" KWHT " sec. time.\n", t1 - t0);

	double r0 = now();
	// InitRadiosity's report (rad.cpp:339-355), on stdout like the reference
	printf("[" KGRN "INFO" KWHT "] Max Patch Area = " KMAG "%f\n" KWHT, radk_model_max_area(model));
	printf("[" KGRN "INFO" KWHT "] Min Patch Area = " KMAG "%f\n" KWHT, radk_model_min_area(model));
	printf("[" KGRN "INFO" KWHT "] Config Table\n");
	printf("       -debug          %i\n", o.debug);
	printf("       -adapt_area     %f\n", o.adapt_area);
	printf("       -sample_area    %f\n", radk_model_sample_area(model));
	printf("       -converge       %f\n", o.converge);
	printf("       -delta_ff       %f\n", o.delta_ff);
	printf("       -write_cycle    %d\n", o.write_cycle);
	printf("       -triangle       %d\n", o.triangle);
	printf("       -light          %f\n", o.light);
	printf("       -interactive    %d\n", o.interactive);
	printf("       -jobs           %d\n", o.jobs);
	printf("       -parallel_src   %d\n", o.parallel_src);
	printf("[" KGRN "INFO" KWHT "] Complete initialization\n");
	fflush(stdout);
	if (!o.split) {
		fprintf(stderr, "[" KRED "DEBUG" KWHT "] Before PrePartitionTriangle %d\n", sc.n0);
		fprintf(stderr, "[" KRED "DEBUG" KWHT "] After PrePartitionTriangle %d\n", sc.trinum);
	}

	radk_opts co;
	radk_default_opts(&co);
	co.converge = o.converge, co.parallel_src = o.parallel_src, co.tri_limit = o.triangle;
	co.area_limit = o.adapt_area, co.sample_area = radk_model_sample_area(model), co.delta_ff = o.delta_ff;
	co.device = o.gpu, co.collect_stats = o.stats, co.split_enabled = o.split;
	if (o.max_tri > 0) co.max_tri = o.max_tri;
	const int ngpu = (o.gpus > 1 && !o.split && !o.stats && o.parallel_src <= 1) ? o.gpus : 1;
	std::vector<radk_ctx *> ctxs((size_t)ngpu, nullptr);
	for (int r = 0; r < ngpu; r++) {
		co.device = o.gpu + r, co.rank = r, co.world = ngpu;
		if (radk_ctx_create(&sc, &co, &ctxs[r]) != RADK_OK) {
			fprintf(stderr, "[" KRED "ERR" KWHT "] %s\n", radk_last_error());
			return 1;
		}
	}
	for (int r = 0; r < ngpu && ngpu > 1; r++)
		if (radk_ctx_peer_connect_local(ctxs[r], ctxs.data(), ngpu) != RADK_OK) {
			fprintf(stderr, "[" KRED "ERR" KWHT "] %s\n", radk_last_error());
			return 1;
		}
	radk_ctx *ctx = ctxs[0];
	// all ranks step together (each blocks on its own stream while exchanging candidates with its peers)
	std::vector<std::string> step_err((size_t)ngpu);
	int failed_rank = -1;
	auto step_all = [&](long long chunk, long long *done, int *converged) {
		if (ngpu == 1) return radk_ctx_step(ctx, chunk, done, converged);
		std::vector<int> rc((size_t)ngpu, RADK_OK);
		std::vector<long long> d((size_t)ngpu, 0);
		std::vector<int> cv((size_t)ngpu, 0);
		std::vector<std::thread> th;
		// radk_last_error() is per thread: each worker keeps its own message
		for (int r = 0; r < ngpu; r++)
			th.emplace_back([&, r] {
				rc[r] = radk_ctx_step(ctxs[r], chunk, &d[r], &cv[r]);
				if (rc[r] != RADK_OK) step_err[r] = radk_last_error();
			});
		for (auto &t : th) t.join();
		for (int r = 0; r < ngpu; r++)
			if (rc[r] != RADK_OK) { failed_rank = r; return rc[r]; }
		*done = d[0], *converged = cv[0];
		return (int)RADK_OK;
	};

	// DoRadiosity's loop (rad.cpp:497-516): iteration `loop` is shot number `loop`; a snapshot is
	// written after every write_cycle-th shot; the iteration that finds no shooter ends the loop.
	const long long W = o.write_cycle > 0 ? o.write_cycle : 1;
	long long shots = 0;
	for (;;) {
		const long long chunk = W - shots % W;
		long long done = 0;
		int converged = 0;
		if (step_all(chunk, &done, &converged) != RADK_OK) {
			if (failed_rank >= 0) fprintf(stderr, "[" KRED "ERR" KWHT "] rank %d: %s\n", failed_rank, step_err[failed_rank].c_str());
			else fprintf(stderr, "[" KRED "ERR" KWHT "] %s\n", radk_last_error());
			return 1;
		}
		shots += done;
		if (converged) {
			if ((shots + 1) % W == 0) fprintf(stderr, "[" KGRN "INFO" KWHT "] Loop %lld BEGIN\n", shots + 1);
			break;
		}
		if (shots % W == 0) {
			fprintf(stderr, "[" KGRN "INFO" KWHT "] Loop %lld BEGIN\n", shots);
			if (!write_snapshot(ctxs, model, o, (int)shots)) return 1;
			fprintf(stderr, "[" KGRN "INFO" KWHT "] END Loop. There are %i triangles.\n", radk_ctx_trinum(ctx));
		}
	}
	fprintf(stderr, "[" KGRN "INFO" KWHT "] Complete in #Iterations %lld\n", shots + 1);
	double w0 = now();
	write_snapshot(ctxs, model, o, 0);
	double w1 = now();
	fprintf(stderr, "[" KGRN "INFO" KWHT "] Store file took " KMAG "%f" KWHT " sec. time.\n", w1 - w0);
	if (o.stats) {
		radk_stats st;
		if (radk_ctx_stats(ctx, &st) == RADK_OK)
			fprintf(stderr, "[" KGRN "INFO" KWHT "] device: %lld shots, %lld reference-equivalent rays (%lld visible), %lld traced, "
				"%.3f ms in the loop, %lld kernel launches\n", st.shots, st.rays, st.visible, st.traced_rays,
				radk_ctx_device_ms(ctx), st.kernel_launches);
	}
	for (radk_ctx *c : ctxs) radk_ctx_destroy(c);
	radk_model_free(model);
	fprintf(stderr, "[" KGRN "INFO" KWHT "] Radiosity Method took " KMAG "%f" KWHT " sec. time.\n", now() - r0);
	return 0;
}
