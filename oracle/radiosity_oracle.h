/* TEST INFRASTRUCTURE - CPU oracle for the progressive-refinement shooting loop.
 *
 * This is a plain-C restatement of the reference's algorithm for the hot path
 * (hw-radiosity src/rad.cpp:387-451, src/shade.cpp:76-168,244-306,435-525,
 * src/raycast.cpp:235-360).  It exists ONLY to check the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.  The
 * product (hw-radiosity_b200/) never links, imports or calls anything here.
 *
 * Parity status: PINNED - checked bit-for-bit against the unmodified reference
 * compiled by oracle/ref_harness (-ffp-contract=off build) on box/room/church/
 * blocks/hall; see tests/test_oracle_pinning.py and tests/golden/.
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off -fopenmp -fPIC -shared (see oracle/Makefile).
 * -ffp-contract=off is REQUIRED: the reference is only deterministic without FMA
 * contraction (SURVEY.md Fact 5).
 */
#ifndef RADIOSITY_ORACLE_H
#define RADIOSITY_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order and sizes as the reference `Triangle` (include/rad.h:25-39), 236 bytes,
 * so that raw dumps of the reference's TriStore can be used directly. */
typedef struct {
	float p[3][3];
	int Frgb[3];
	int Brgb[3];
	float accB[3][3];
	float deltaaccB[3][3];
	float deltaB[3];
	float se[3][4];
	float c[3], n[3];
	float area;
	int parent;
	int neighbor[3];
} orc_triangle;

/* Same layout as the reference `TreeNode` (include/rad.h:41-45), 60 bytes. */
typedef struct {
	float g0[3], g1[3];
	int sub[8];
	int list;
} orc_node;

typedef struct {
	long long rays;		/* RayHitted calls (== CalFF calls surviving the culls) */
	long long visible;	/* rays whose result == the shooter's logical index */
	long long calff;	/* CalFF calls */
	long long leaf_visits;	/* TriListHitted calls */
	long long descent_steps;/* TreeNodeNum inner-loop iterations */
	long long tri_tests;	/* TriHitted calls */
	long long tri_accepts;	/* TriHitted returning true */
	long long precision_exceeded;	/* raycast.cpp:344-351 fallback taken */
	long long receivers;	/* receivers that were shaded (not skipped) */
	long long splits;	/* successful centre splits (variant A only) */
} orc_stats;

typedef struct {
	orc_triangle *tri;	/* TriStore */
	int n0;			/* input triangle count; occluders are tri[0..n0) */
	int trinum;
	int capacity;		/* slots allocated in tri[] (>= trinum; matters for splitting) */
	const orc_node *nodes;
	int nnodes;
	const int *list;
	int nlist;
	float G0[3], G1[3];
	float mingridlen;
	int gridnum;
	/* options (rad.cpp:17-31) */
	float converge;		/* ConvergeLimit */
	int parallel_src;	/* ParallelSource */
	int tri_limit;		/* TriangleLimit */
	float area_limit;	/* AreaLimit */
	float sample_area;	/* SampleArea */
	float delta_ff;		/* DeltaFFLimit */
	int max_tri;		/* the reference's MaxTri (rad.h:9); 1000000 there */
	int split_enabled;	/* 0: shipped config (DONT_PARTITION); 1: "variant A" centre split */
	int threads;		/* OpenMP threads for the receiver loop (frozen mesh only) */
	orc_stats stats;
} orc_scene;

/* rad.cpp:387-404 listCanditate.  Writes up to `cap` shooter slots, returns the count. */
int orc_select(const orc_scene *s, int *out, int cap);

/* rad.cpp:406-451 parallelRadiosity: one shot.  Returns 0 when converged, 1 after a shot.
 * If shooters_out != NULL the chosen shooter slots are written there (up to cap). */
int orc_shot(orc_scene *s, int *shooters_out, int cap, int *nshooters);

/* One shot with a caller-chosen list of shooters (same receiver loop, same zeroing). */
void orc_shot_with(orc_scene *s, const int *shooters, int nshooters);

/* raycast.cpp:335-360 RayHitted. */
int orc_ray_hitted(orc_scene *s, const float p[3], const float v[3], int from, int to);

/* batch form for tests */
void orc_ray_hitted_batch(orc_scene *s, const float *p, const float *v, const int *from,
			  int n, int *out);

/* shade.cpp:141-168 AdaptCalFF(CalFF(centre)) for one receiver vertex, returning the live value */
float orc_vertex_ff(orc_scene *s, int shooter, int receiver, const float p[3]);

/* rad.cpp:167-178 snapshot colour of one slot */
void orc_snapshot_color(const orc_triangle *t, int rgb[3]);

#ifdef __cplusplus
}
#endif
#endif
