/* radk.h - C ABI of the B200-native progressive-refinement shooting loop.
 *
 * Drop-in boundary for hw-radiosity's hot path.  The reference has no FFI layer; its seam is
 * three C++ functions plus global arrays:
 *     void BuildTree(void);  int RayHitted(Vector p, Vector v, int otri, int ttri);   include/raycast.h:4-5
 *     int  Shade(TrianglePtr, int, TrianglePtr, int, int);  void PrePartitionTriangles();  include/shade.h:4-5
 *     static int parallelRadiosity();                                                 src/rad.cpp:406
 * This header replaces that seam with plain-C entry points (pointers and sizes only, caller-owned
 * host buffers, 0 = success, no exceptions, no callbacks).  A context is not thread-safe.
 *
 * All compute entry points run on the GPU (sm_100a).  There is NO CPU fallback: without a usable
 * CUDA device every radk_ctx_* / radk_ray_* call fails with RADK_ERR_CUDA.
 */
#ifndef RADK_H
#define RADK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RADK_VERSION 1

enum {
	RADK_OK = 0,
	RADK_ERR_ARG = 1,	/* bad argument */
	RADK_ERR_IO = 2,	/* file could not be opened / parsed (reference: "File open error", rad.cpp:641-644) */
	RADK_ERR_LIMIT = 3,	/* a reference limit was exceeded (reference: exit(1), raycast.cpp:175-178) */
	RADK_ERR_CUDA = 4,	/* CUDA error or no device */
	RADK_ERR_STATE = 5	/* call not valid in the current state */
};

/* ---- records -------------------------------------------------------------------------------
 * radk_triangle has the field order and sizes of the reference `Triangle` (include/rad.h:25-39,
 * 236 bytes); radk_tree_node those of `TreeNode` (include/rad.h:41-45, 60 bytes).  A maintainer
 * can therefore pass &TriStore[0] and &TreeNodeStore[0] straight through. */
typedef struct radk_triangle {
	float p[3][3];		/* vertices */
	int Frgb[3];		/* reflectance colour */
	int Brgb[3];		/* emission colour; light iff any component == 255 (rad.h:47-49) */
	float accB[3][3];	/* accumulated radiosity per vertex, RGB */
	float deltaaccB[3][3];	/* last increment (only read by the dormant neighbour un-shade) */
	float deltaB[3];	/* unshot radiosity, RGB */
	float se[3][4];		/* side plane equations */
	float c[3], n[3];	/* centre, normal */
	float area;
	int parent;		/* original triangle this slot descends from; -1 = logical (not a receiver) */
	int neighbor[3];
} radk_triangle;

typedef struct radk_tree_node {
	float g0[3], g1[3];
	int sub[8];
	int list;		/* >= 0: leaf, offset of its -1 terminated index list */
} radk_tree_node;

/* Everything the loop reads: the reference's globals TriStore/trinum/G0/G1 (rad.cpp:33-36) and the
 * file-statics of raycast.cpp:9-16.  Pointers are borrowed for the duration of the call only. */
typedef struct radk_scene {
	const radk_triangle *tri;	/* TriStore[0..trinum) AFTER PrePartitionTriangles() */
	int n0;				/* input triangle count: occluders are tri[0..n0) */
	int trinum;
	const radk_tree_node *nodes;	/* TreeNodeStore[0..nnodes), root = 0 */
	int nnodes;
	const int *list;		/* TriListStore[0..nlist) */
	int nlist;
	float G0[3], G1[3];		/* scene cube */
	float min_grid_len;		/* MinGridLen */
	int grid_num;			/* GridNum */
} radk_scene;

/* Option globals of rad.cpp:17-31 that the loop reads. */
typedef struct radk_opts {
	float converge;		/* -converge     ConvergeLimit  (default 1200) */
	int parallel_src;	/* -parallel_src ParallelSource (default 1) */
	int tri_limit;		/* -triangle     TriangleLimit  (default 1000000) */
	float area_limit;	/* -adapt_area   AreaLimit      (default 5)      split path only */
	float sample_area;	/* -sample_area  SampleArea     (<0: model dependent) split path only */
	float delta_ff;		/* -delta_ff     DeltaFFLimit   (default 5e-6)   split path only */
	int max_tri;		/* the reference's MaxTri (rad.h:9); 0 = 1000000 */
	int split_enabled;	/* 0 = shipped configuration (DONT_PARTITION); 1 = adaptive centre split */
	int device;		/* CUDA device ordinal */
	int rank, world;	/* receiver shard: the sorted receivers are dealt to the ranks in blocks of 512 */
	int collect_stats;	/* 1: count rays / visits / tests on the device (slower kernels) */
	int dedup;		/* 1 (default when 0 is not forced by split_enabled): evaluate each unique
				 * (vertex, parent) once per shot; -1: force per-receiver evaluation */
	int per_shot_kernels;	/* 1: run every shot as its own kernel pair (K3a, K3b) instead of the batched
				 * predict / form-factor-cache / persistent-loop path (same results) */
	int ff_rows;		/* form-factor cache rows (shooters evaluated per batch); 0 = default 32 */
	int reserved[4];
} radk_opts;

typedef struct radk_stats {
	long long shots;
	long long rays;			/* reference-equivalent RayHitted calls */
	long long visible;
	long long calff;		/* reference-equivalent CalFF calls */
	long long traced_rays;		/* rays actually walked through the octree on the device */
	long long leaf_visits, descent_steps, tri_tests, tri_accepts;	/* per traced ray */
	long long precision_exceeded;
	long long receivers;		/* shaded receivers summed over shots */
	long long splits;
	long long kernel_launches;	/* kernels this context has launched */
	long long ray_guard_trips;	/* walks aborted by the device iteration guard (reference would spin) */
	/* batched path (since the last create / reset / rewind): */
	long long ff_rows_filled;	/* form-factor cache rows computed (one per predicted shooter) */
	long long rounds;		/* predict -> form factors -> chain -> apply -> commit rounds */
	long long shots_scheduled;	/* shots the chains scheduled; shots - of those - that were committed is `shots` */
	long long rounds_reapplied;	/* rounds whose schedule went past the first violation (valid prefix applied again) */
} radk_stats;

void radk_default_opts(radk_opts *o);
const char *radk_last_error(void);
int radk_version(void);

/* ---- host front end (init time; replaces InitModelFromFile rad.cpp:244-252, InitRadiosity
 *      rad.cpp:259-356, PrePartitionTriangles shade.cpp:556-598, BuildTree raycast.cpp:216-233) -- */
typedef struct radk_model radk_model;

typedef struct radk_model_opts {
	float light_scale;	/* -light (default 1) */
	int tri_limit;		/* -triangle */
	int max_tri;		/* 0 = 1000000 */
	int max_tree_nodes;	/* 0 = 20000 (rad.h:11) */
	int prepartition;	/* 1 = run PrePartitionTriangles as the reference does (shipped config); 0 = skip;
				 * 2 = exhaustive variant for synthetic scaling scenes: stale work-set entries and
				 *     refused splits are skipped instead of ending the refinement (NOT reference behaviour) */
	float sample_area;	/* <0: maxarea/100 (rad.cpp:337-338) */
	int reserved[6];
} radk_model_opts;

void radk_default_model_opts(radk_model_opts *o);
/* Reads a .tri file (rad.cpp:47-111), builds the octree, initialises emitters and neighbours and
 * (optionally) pre-partitions.  *out owns all storage. */
int radk_model_load(const char *tri_path, const radk_model_opts *o, radk_model **out);
/* Same, from an in-memory .tri text. */
int radk_model_parse(const char *text, size_t len, const radk_model_opts *o, radk_model **out);
void radk_model_free(radk_model *m);
/* Borrowed view valid until the model is freed or mutated. */
int radk_model_scene(const radk_model *m, radk_scene *out);
float radk_model_sample_area(const radk_model *m);
float radk_model_min_area(const radk_model *m);
float radk_model_max_area(const radk_model *m);
/* Mutable TriStore access (to merge downloaded radiosity back before writing). */
radk_triangle *radk_model_triangles(radk_model *m, int *trinum);
/* Replace the model's slots by `trinum` records (after an adaptive-split run grew the pool). */
int radk_model_set_triangles(radk_model *m, const radk_triangle *tri, int trinum);
/* Raw dump in the harness "RADSCN1" layout (tests compare it with the reference's dump). */
int radk_model_dump(const radk_model *m, const radk_opts *o, const char *path);
/* PrintOut (rad.cpp:117-242) body: "<prefix>.fin" when loop == 0 else "<prefix>.<loop>". */
int radk_model_write(const radk_model *m, const char *prefix, int loop);
/* Snapshot writer for a frozen mesh: same file as radk_model_write, but from per-slot
 * colours rgb[trinum][3] (radk_ctx_snapshot_colors) instead of the model's accB; the vertex/normal text is
 * formatted once and reused, and the file is written by a background thread (at most one in flight; the
 * next call, radk_model_write_wait or radk_model_free joins it).  Returns the previous write's status. */
int radk_model_write_colors(radk_model *m, const char *prefix, int loop, const int *rgb, int trinum);
int radk_model_write_wait(radk_model *m);
/* The same text into a caller buffer; returns the needed size in *len. */
int radk_model_format(const radk_model *m, char *buf, size_t cap, size_t *len);

/* ---- device context (replaces the loop: listCanditate + parallelRadiosity + Shade + RayHitted) ---- */
typedef struct radk_ctx radk_ctx;

/* Uploads the scene (receivers = slots with parent >= 0, occluders = tri[0..n0), octree). */
int radk_ctx_create(const radk_scene *s, const radk_opts *o, radk_ctx **out);
void radk_ctx_destroy(radk_ctx *c);
/* Re-upload radiosity state (accB, deltaB) from the scene's slots; geometry unchanged. */
int radk_ctx_reset(radk_ctx *c, const radk_scene *s);

/* Restore the radiosity state of the last create/reset from its device-resident copy (no host
 * traffic); used to re-run the same scene. */
int radk_ctx_rewind(radk_ctx *c);

/* Runs up to max_shots iterations of parallelRadiosity() (rad.cpp:406-451) on the device.
 * *converged = 1 when listCanditate() came back empty (that iteration is not counted as a shot).
 * max_shots < 0: run to convergence. */
int radk_ctx_step(radk_ctx *c, long long max_shots, long long *shots_done, int *converged);
/* listCanditate() (rad.cpp:387-404) on the current state: slot indices, descending (delta, idx). */
int radk_ctx_select(radk_ctx *c, int *slots, int cap, int *n);
/* One pass with caller-chosen shooters (same receiver rules, zeroes their deltaB afterwards). */
int radk_ctx_shoot(radk_ctx *c, const int *slots, int n);
/* Shooter slots of the shots run so far by radk_ctx_step (first parallel source of each), newest
 * `cap` at most; returns total count in *n. */
int radk_ctx_shooter_log(radk_ctx *c, int *slots, int cap, long long *n);

int radk_ctx_trinum(const radk_ctx *c);
/* Download per-slot state into caller arrays indexed by slot (any pointer may be NULL):
 * accB[trinum][3][3], deltaB[trinum][3], parent[trinum]. Logical slots are left untouched. */
int radk_ctx_download(radk_ctx *c, float *accB, float *deltaB, int *parent, int trinum);
/* Download whole records (split runs): writes trinum radk_triangle records. */
int radk_ctx_download_triangles(radk_ctx *c, radk_triangle *tri, int cap, int *trinum);
/* Snapshot colours (rad.cpp:167-178) per slot: rgb[trinum][3]; logical slots get -1. */
int radk_ctx_snapshot_colors(radk_ctx *c, int *rgb, int trinum);
int radk_ctx_stats(radk_ctx *c, radk_stats *st);
/* Device time (ms, CUDA events on the context's stream) spent inside radk_ctx_step so far. */
double radk_ctx_device_ms(const radk_ctx *c);
/* Diagnostics (contexts created with RADK_DEBUG_PREDICT=1 in the environment; -1 everywhere otherwise): mean time in us,
 * after the start of a round of the batched loop, at which each of its kernels started / ended, over the last <= 64
 * rounds - out[0..10] = predict start, predict end, merge start, merge end, candidates' rays start, form factors start,
 * chain start, chain end, apply start, commit start, commit's bookkeeping done; out[11] = median gap to the next round;
 * out[12] = rounds averaged; out[13] = mismatches of the top-64 selection self-check.  No counterpart in the reference. */
int radk_ctx_round_timeline(radk_ctx *c, double out[14]);
/* Profiling aid: runs up to max_shots shots (max_shots < 0: to convergence) with a CUDA-event pair
 * around every kernel launch on the context's stream and accumulates device time per kernel class:
 *   ms[0]/n[0] selection   (predict_kernel, or select_kernel on the per-shot path)
 *   ms[1]/n[1] visibility + form factors (ff_batch_kernel, or vertex_ff_kernel)   <- the traversal
 *   ms[2]/n[2] radiosity update + argmax (shots_kernel, or update_select_kernel)
 * n[] = launches per class; *shots = shots run. */
int radk_ctx_profile(radk_ctx *c, long long max_shots, double ms[3], long long n[3], long long *shots);
/* Sizes of this context's shard: receivers, unique vertices, occluders, octree nodes. */
int radk_ctx_sizes(const radk_ctx *c, int *receivers, int *unique_vertices, int *occluders, int *nodes);

/* Batch RayHitted (raycast.cpp:335-360): n rays (p[n][3], v[n][3], from[n]) -> hit[n]. */
int radk_ray_hitted(radk_ctx *c, const float *p, const float *v, const int *from, int n, int *hit);
/* The question CalFF asks of RayHitted (shade.cpp:124): does the ray return `logsrc`?  sees[i] = 1 iff
 * RayHitted(p[i], v[i], from[i], logsrc) == logsrc, answered by the walk the batched form-factor kernel runs
 * (bounded any-hit leaf visits with the per-shooter leaf mask) instead of the nearest-hit walk of radk_ray_hitted. */
int radk_ray_sees(radk_ctx *c, const float *p, const float *v, const int *from, int logsrc, int n, int *sees);
/* Batch of AdaptCalFF(CalFF(..)) (shade.cpp:442-446) for n (receiver slot, vertex 0..2) pairs
 * against shooter slot `shooter`. */
int radk_vertex_ff(radk_ctx *c, int shooter, const int *receiver, const int *vertex, int n, float *ff);

/* ---- multi-GPU: one process per GPU, receivers sharded by opts.rank / opts.world -------------------
 * Octree, occluders and per-original tables are replicated; each context owns a contiguous range of the
 * sorted receivers.  The only exchange on the path is the shooter selection:
 * per round the ranks' top candidates (with complete shooter records), per shot the ranks' best
 * (unshot energy, slot).  Both go through peer-mapped mailboxes written with plain NVLink stores from
 * inside the kernels; the host only has to introduce the ranks to each other once:
 *   1. every rank calls radk_ctx_peer_handle and the launcher all-gathers the 64-byte handles
 *      (torch.distributed / NCCL in hw-radiosity_b200/dist.py),
 *   2. every rank calls radk_ctx_peer_connect with the world x 64 bytes.
 * After that radk_ctx_step must be called by all ranks with the same arguments; every rank ends with
 * the same shooter sequence and its own shard of the radiosity, bit-identical to a single-GPU run.
 * reset / rewind must be called on all ranks alike.  A peer that stops responding makes radk_ctx_step
 * fail with RADK_ERR_STATE after a device-side timeout instead of hanging. */
#define RADK_PEER_HANDLE_BYTES 64
int radk_ctx_peer_handle(radk_ctx *c, void *handle);
int radk_ctx_peer_connect(radk_ctx *c, const void *handles, int world);
/* Same-process variant (tests, single-process multi-GPU drivers): peers[r] is rank r's context. */
int radk_ctx_peer_connect_local(radk_ctx *c, radk_ctx *const *peers, int world);
/* Receivers owned by this context: slots[] gets their TriStore slot indices (cap entries at most). */
int radk_ctx_owned_slots(radk_ctx *c, int *slots, int cap, int *n);

#ifdef __cplusplus
}
#endif
#endif
