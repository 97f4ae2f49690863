// TEST INFRASTRUCTURE (oracle/_ref harness). Not part of the product.
// Translation unit 3 of 3: the UNMODIFIED reference shading code
// (/root/reference/src/shade.cpp) with its calls to RayHitted (shade.cpp:124)
// routed through a counting / logging hook.  The hook calls the real RayHitted
// (raycast.cpp:335) and returns its result unchanged.
#include <stdio.h>
#include <string.h>
#include <omp.h>
#include REF_VECTOR_H
#include "harness_api.h"

int RayHitted(Vector p, Vector v, int otri, int ttri);	// the real one (raycast.h:5)

#define HARNESS_MAX_THREADS 512
static long long g_cnt[HARNESS_MAX_THREADS][8];		// [thread][0]=rays [1]=visible, padded to a cache line
static FILE *g_raylog = NULL;

void harness_ray_log_open(const char *path) { g_raylog = fopen(path, "wb"); }
void harness_ray_log_close(void) { if (g_raylog) { fclose(g_raylog); g_raylog = NULL; } }
void harness_ray_counters_reset(void) { memset(g_cnt, 0, sizeof(g_cnt)); }
void harness_ray_counters(long long *rays, long long *visible)
{
	*rays = *visible = 0;
	for (int i = 0; i < HARNESS_MAX_THREADS; i++) *rays += g_cnt[i][0], *visible += g_cnt[i][1];
}

static inline int RayHitted_hook(Vector p, Vector v, int otri, int ttri)
{
	int r = RayHitted(p, v, otri, ttri);
	int t = omp_get_thread_num();
	g_cnt[t][0]++;
	g_cnt[t][1] += (r == ttri);
	if (g_raylog) {
		struct { float p[3], v[3]; int from, to, result; } rec;
		for (int i = 0; i < 3; i++) rec.p[i] = p[i], rec.v[i] = v[i];
		rec.from = otri, rec.to = ttri, rec.result = r;
		fwrite(&rec, sizeof(rec), 1, g_raylog);
	}
	return r;
}

#define RayHitted RayHitted_hook
#define _RAYCAST_H		/* suppress include/raycast.h: declarations supplied here */
void BuildTree(void);
#include REF_SHADE_CPP
