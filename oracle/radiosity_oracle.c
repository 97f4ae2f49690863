/* TEST INFRASTRUCTURE - CPU oracle (see radiosity_oracle.h for scope and rules).
 *
 * Every function names the reference lines it restates.  Arithmetic follows the
 * reference expression by expression (float unless the reference's C++ promotes to
 * double through a double literal), evaluated left to right, no FMA contraction.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "radiosity_oracle.h"

#define ORC_PI (3.1415926535897932)	/* rad.h:5 */

/* vector.h:36 InnerProd: (a0*b0 + a1*b1) + a2*b2 */
static inline float dot3(const float *a, const float *b)
{
	return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
/* vector.h:68-70 */
static inline float norm2f(const float *v) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }
/* vector.h:65-67; (float)sqrt((double)x) == sqrtf(x) for IEEE binary32/64 */
static inline float normf(const float *v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

/* ------------------------------------------------------------------ visibility */

/* raycast.cpp:254-281 TreeNodeNum */
static int tree_node_num(const orc_scene *s, const float *q, orc_stats *st)
{
	if (q[0] < s->G0[0] || q[0] > s->G1[0] ||
	    q[1] < s->G0[1] || q[1] > s->G1[1] ||
	    q[2] < s->G0[2] || q[2] > s->G1[2])
		return -1;
	int gnum[3];
	for (int i = 0; i < 3; i++)
		gnum[i] = (int)(((double)q[i] - (double)s->G0[i]) / (double)s->mingridlen);
	int ctn = 0;
	int gsize = s->gridnum;
	while (s->nodes[ctn].list < 0) {
		gsize >>= 1;
		int index = 0;
		for (int i = 0; i < 3; i++) {
			if (gnum[i] >= gsize) {
				index |= (1 << i);
				gnum[i] -= gsize;
			}
		}
		ctn = s->nodes[ctn].sub[index];
		st->descent_steps++;
	}
	return ctn;
}

/* raycast.cpp:235-251 NextPoint */
static float next_point(const float *p, const float *v, const float *g0, const float *g1, float *q)
{
	float t_min = 1000000.0f;
	for (int i = 0; i < 3; i++) {
		if (fabs((double)v[i]) > 1e-6) {
			float s;
			if (v[i] > 0)
				s = (g1[i] - p[i]) / v[i];
			else
				s = (g0[i] - p[i]) / v[i];
			t_min = (s < t_min) ? s : t_min;	/* std::min(t_min, s) */
		}
	}
	t_min = (float)((double)t_min + 0.1);
	q[0] = p[0] + t_min * v[0];
	q[1] = p[1] + t_min * v[1];
	q[2] = p[2] + t_min * v[2];
	return t_min;
}

/* raycast.cpp:283-313 TriHitted */
static int tri_hitted(const float *p, const float *v, const orc_triangle *tp, float *t_min)
{
	float a1 = dot3(v, tp->n);
	float vv[3] = { tp->c[0] - p[0], tp->c[1] - p[1], tp->c[2] - p[2] };
	float a2 = dot3(vv, tp->n);
	float t = a2 / a1;
	if ((double)t < 0.0001)
		return 0;
	if (t >= *t_min)
		return 0;
	float q[3] = { p[0] + t * v[0], p[1] + t * v[1], p[2] + t * v[2] };
	for (int k = 0; k < 3; k++)
		if ((q[0] * tp->se[k][0] + q[1] * tp->se[k][1] + q[2] * tp->se[k][2] + tp->se[k][3]) < 0.0)
			return 0;
	*t_min = t;
	return 1;
}

/* raycast.cpp:316-333 TriListHitted */
static int tri_list_hitted(const orc_scene *s, const float *p, const float *v, int store, int ban,
			   orc_stats *st)
{
	if (store < 0)
		return -1;
	int hit = -1, idx;
	float t_min = 1000000.0f;
	while ((idx = s->list[store++]) >= 0) {
		if (idx == ban)
			continue;
		st->tri_tests++;
		if (tri_hitted(p, v, &s->tri[idx], &t_min)) {
			st->tri_accepts++;
			hit = idx;
		}
	}
	return hit;
}

/* raycast.cpp:335-360 RayHitted.  `to` is unused by the walk itself (as in the reference). */
static int ray_hitted(const orc_scene *s, const float *p, const float *v, int from, int to, orc_stats *st)
{
	float t = 0.f;
	int prev = -1, cur, hit;
	float q[3] = { p[0], p[1], p[2] };
	long guard = 0, steps = 0;
	(void)to;
	while ((cur = tree_node_num(s, q, st)) >= 0) {
		/* Not in the reference: when GridNum is not a power of two its descent disagrees with the leaf boxes
		 * and this loop can bounce between two leaves forever; cut like the device does (rt_core.cuh kMaxWalkSteps). */
		if (++steps > 1024)
			return -1;
		if (prev == cur) {
			st->precision_exceeded++;
			t = (float)((double)t + 1.0);
			q[0] = p[0] + t * v[0];
			q[1] = p[1] + t * v[1];
			q[2] = p[2] + t * v[2];
			if (++guard > 100000)	/* the reference would spin forever here */
				return -1;
			continue;
		}
		const orc_node *nd = &s->nodes[cur];
		st->leaf_visits++;
		if ((hit = tri_list_hitted(s, p, v, nd->list, from, st)) >= 0)
			return hit;
		prev = cur;
		t = next_point(p, v, nd->g0, nd->g1, q);
	}
	return -1;
}

int orc_ray_hitted(orc_scene *s, const float p[3], const float v[3], int from, int to)
{
	return ray_hitted(s, p, v, from, to, &s->stats);
}

void orc_ray_hitted_batch(orc_scene *s, const float *p, const float *v, const int *from, int n, int *out)
{
	for (int i = 0; i < n; i++)
		out[i] = ray_hitted(s, p + 3 * i, v + 3 * i, from[i], -1, &s->stats);
}

/* ------------------------------------------------------------------ form factor */

/* the three fields CalFF reads from its source triangle */
typedef struct { float c[3], n[3], area; } ff_sample;

/* shade.cpp:107-127 CalFF */
static float cal_ff(const orc_scene *s, const ff_sample *src, int logsrc, const float *ndes, int logdest,
		    const float *p, orc_stats *st)
{
	st->calff++;
	float dir[3] = { src->c[0] - p[0], src->c[1] - p[1], src->c[2] - p[2] };
	float dotSrc = dot3(dir, src->n);
	float dotDes = -dot3(dir, ndes);
	if ((dotSrc < 0) != (dotDes < 0))
		return 0.0f;
	float dirLen = normf(dir), srcLen = normf(src->n), desLen = normf(ndes);
	float ctheta1 = dotSrc / dirLen / srcLen;
	float ctheta2 = dotDes / dirLen / desLen;
	float ff = ctheta1 * ctheta2;
	if (ff <= 0.0)
		return 0.0f;
	ff = (float)((double)ff * ((double)src->area / ((double)norm2f(dir) * ORC_PI)));
	if (ff <= 0.0)
		return 0.0f;
	st->rays++;
	if (ray_hitted(s, p, dir, logdest, logsrc, st) == logsrc) {
		st->visible++;
		return ff;
	}
	return 0.0f;
}

/* vector.h:54-57 CalCenter: float sum, double divide by 3.0, rounded to float */
static void cal_center(const float p[3][3], float *c)
{
	for (int k = 0; k < 3; k++)
		c[k] = (float)((double)(p[0][k] + p[1][k] + p[2][k]) / 3.0);
}

/* shade.cpp:146-156 + 76-99: longest edge (first max wins) and the two half samples.
 * Returns 0 when the shooter is too small to be split (maxlength < 0.1). */
static int shooter_halves(const orc_triangle *src, ff_sample *t1, ff_sample *t2)
{
	int edge = 0;
	float maxlength = 0.0f;
	for (int v = 0; v < 3; v++) {
		const float *a = src->p[v], *b = src->p[(v + 1) % 3];
		float l[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] };
		float length = norm2f(l);
		if (maxlength < length)
			maxlength = length, edge = v;
	}
	if ((double)maxlength < 0.1)
		return 0;
	int i0 = edge, i1 = (edge + 1) % 3, i2 = (edge + 2) % 3;
	float m[3], tp[3][3];
	for (int k = 0; k < 3; k++)
		m[k] = (float)((double)(src->p[i0][k] + src->p[i1][k]) / 2.0);	/* vector.h:27-30 */
	for (int k = 0; k < 3; k++) tp[0][k] = m[k], tp[1][k] = src->p[i2][k], tp[2][k] = src->p[i0][k];
	cal_center(tp, t1->c);
	for (int k = 0; k < 3; k++) tp[0][k] = m[k], tp[1][k] = src->p[i1][k], tp[2][k] = src->p[i2][k];
	cal_center(tp, t2->c);
	for (int k = 0; k < 3; k++) t1->n[k] = t2->n[k] = src->n[k];
	t1->area = t2->area = (float)((double)src->area / 2.0);
	return 1;
}

/* shade.cpp:141-168 AdaptCalFF(CalFF(centre), ...) as it executes: because of the stray ';'
 * at shade.cpp:164 the result is always CalFF(t1)+CalFF(t2) once the shooter is splittable;
 * the centre sample is still evaluated first (its ray is cast and counted). */
static float adapt_ff(const orc_scene *s, const orc_triangle *src, int logsrc, const float *ndes, int logdest,
		      const float *p, orc_stats *st)
{
	ff_sample c0, t1, t2;
	memcpy(c0.c, src->c, sizeof(c0.c));
	memcpy(c0.n, src->n, sizeof(c0.n));
	c0.area = src->area;
	float ff = cal_ff(s, &c0, logsrc, ndes, logdest, p, st);
	if (!shooter_halves(src, &t1, &t2))
		return ff;
	float ff1 = cal_ff(s, &t1, logsrc, ndes, logdest, p, st);
	float ff2 = cal_ff(s, &t2, logsrc, ndes, logdest, p, st);
	return ff1 + ff2;
}

float orc_vertex_ff(orc_scene *s, int shooter, int receiver, const float p[3])
{
	const orc_triangle *src = &s->tri[shooter], *des = &s->tri[receiver];
	return adapt_ff(s, src, src->parent, des->n, des->parent, p, &s->stats);
}

/* shade.cpp:244-259 CalRadiosity */
static void cal_radiosity(const orc_triangle *src, orc_triangle *des, const float ff[3])
{
	const float k = (float)(0.5 / 255.0);
	float lo[3];
	for (int c = 0; c < 3; c++)
		lo[c] = des->Frgb[c] * k * (src->deltaB[c]) / 3;
	for (int v = 0; v < 3; v++)
		for (int c = 0; c < 3; c++) {
			float d = lo[c] * ff[v];
			des->deltaB[c] += d;
			des->accB[v][c] += d;
			des->deltaaccB[v][c] = d;
		}
}

/* ------------------------------------------------------------------ centre split (variant A) */

static int pool_limit(const orc_scene *s)
{
	return s->max_tri < s->tri_limit ? s->max_tri : s->tri_limit;
}
/* shade.cpp:26-30 */
static int is_full_pool(const orc_scene *s, int need) { return s->trinum + need >= pool_limit(s); }
/* shade.cpp:14-25 */
static int alloc_triangle(orc_scene *s)
{
	if (s->trinum == s->max_tri || s->trinum == s->tri_limit || s->trinum >= s->capacity)
		return -1;
	return s->trinum++;
}

/* vector.h:58-62 CalArea (normal left un-normalised) */
static void cal_area(orc_triangle *t)
{
	float v0[3], v1[3];
	for (int k = 0; k < 3; k++) v0[k] = t->p[1][k] - t->p[0][k], v1[k] = t->p[2][k] - t->p[0][k];
	t->n[0] = v0[1] * v1[2] - v1[1] * v0[2];
	t->n[1] = v0[2] * v1[0] - v1[2] * v0[0];
	t->n[2] = v0[0] * v1[1] - v1[0] * v0[1];
	t->area = normf(t->n) / 2;
}

/* shade.cpp:261-306 PartitionTriangleWithCenterAndStore + 213-241 PartitionDestination (3-way).
 * Returns -1 refused, 0 original made logical, 1 replaced in place (t3 == src slot). */
static int center_split(orc_scene *s, int real, int *id1, int *id2, int *id3)
{
	orc_triangle *src = &s->tri[real];
	if (is_full_pool(s, 4) || (double)src->area < 1e-8)
		return -1;
	float maxLen = 0.0f, minLen = 1e+30f;
	for (int v = 0; v < 3; v++) {
		const float *a = src->p[v], *b = src->p[(v + 1) % 3];
		float l[3] = { b[0] - a[0], b[1] - a[1], b[2] - a[2] };
		float length = norm2f(l);
		maxLen = (maxLen < length) ? length : maxLen;	/* std::max */
		minLen = (length < minLen) ? length : minLen;	/* std::min */
	}
	if ((double)(minLen / maxLen) < 1e-1 || minLen < 1)
		return -1;
	int a = alloc_triangle(s), b = alloc_triangle(s), c = alloc_triangle(s);
	if (a < 0 || b < 0 || c < 0)
		abort();	/* the reference would index TriStore[-1] here; cannot happen after is_full_pool */
	orc_triangle *t1 = &s->tri[a], *t2 = &s->tri[b], *t3 = &s->tri[c];
	src = &s->tri[real];
	*t1 = *src; *t2 = *src; *t3 = *src;
	memcpy(t1->p[0], src->c, sizeof(src->c));
	cal_center((const float (*)[3])t1->p, t1->c);
	cal_area(t1);
	memcpy(t2->p[1], src->c, sizeof(src->c));
	cal_center((const float (*)[3])t2->p, t2->c);
	cal_area(t2);
	memcpy(t3->p[2], src->c, sizeof(src->c));
	cal_center((const float (*)[3])t3->p, t3->c);
	cal_area(t3);
	t1->neighbor[0] = c, t1->neighbor[2] = b;
	t2->neighbor[1] = a, t2->neighbor[0] = c;
	t3->neighbor[2] = b, t3->neighbor[1] = a;
	t1->parent = t2->parent = t3->parent = src->parent;
	*id1 = a, *id2 = b, *id3 = c;
	if (src->parent == real) {
		src->parent = -1;
		return 0;
	}
	*src = *t3;
	*id3 = real;
	s->trinum--;
	return 1;
}

/* shade.cpp:49-68 GetPointInTriangle (glibc rand(), default seed unless the caller srand()s) */
static float point_in_triangle(const orc_triangle *t, float *p, const float ff[3])
{
	float r1 = (float)rand(), r2 = (float)rand(), r3 = (float)rand();
	float tot = r1 + r2 + r3;
	r1 /= tot; r2 /= tot; r3 /= tot;
	for (int v = 0; v < 3; v++)
		p[v] = r1 * t->p[0][v] + r2 * t->p[1][v] + r3 * t->p[2][v];
	return r1 * ff[0] + r2 * ff[1] + r3 * ff[2];
}

/* ------------------------------------------------------------------ Shade */

/* shade.cpp:435-525 Shade under {ADAPT, PARALLEL} with or without DONT_PARTITION. */
static int shade(orc_scene *s, int srcidx, int logsrc, int desidx, int logdest, orc_stats *st)
{
	const orc_triangle *src = &s->tri[srcidx];
	orc_triangle *des = &s->tri[desidx];
	float ff[3];
	for (int v = 0; v < 3; v++)
		ff[v] = adapt_ff(s, src, logsrc, des->n, logdest, des->p[v], st);
	if (des->area < s->area_limit || !s->split_enabled) {
		cal_radiosity(src, des, ff);
		return 0;
	}
	if (is_full_pool(s, 4) || (double)des->area < 1e-8) {
		cal_radiosity(src, des, ff);
		return 0;
	}
	float groundFF = (float)((double)(ff[0] + ff[1] + ff[2]) / 3.0);
	float ffs = adapt_ff(s, src, logsrc, des->n, logdest, des->c, st);
	float deltaff;
	if ((deltaff = fabsf(groundFF - ffs)) < s->delta_ff) {
		int samplenum = (int)(des->area / s->sample_area);
		for (int v = 0; v < samplenum; v++) {
			float tmp[3];
			groundFF = point_in_triangle(des, tmp, ff);
			ffs = adapt_ff(s, src, logsrc, des->n, logdest, tmp, st);
			if ((deltaff = fabsf(groundFF - ffs)) > s->delta_ff)
				break;
		}
		if (deltaff < s->delta_ff) {
			cal_radiosity(src, des, ff);
			return 0;
		}
	}
	int a, b, c;
	int flag = center_split(s, desidx, &a, &b, &c);
	src = &s->tri[srcidx];
	des = &s->tri[desidx];
	if (flag >= 0)
		st->splits++;
	if (flag == 1) {
		orc_triangle *t3 = &s->tri[c];
		float ff2[3];
		ff2[2] = adapt_ff(s, src, logsrc, t3->n, logdest, t3->p[2], st);
		ff2[0] = ff[0];
		ff2[1] = ff[1];
		cal_radiosity(src, t3, ff2);
	}
	if (flag < 0)
		cal_radiosity(src, des, ff);
	return flag;
}

/* ------------------------------------------------------------------ the loop */

typedef struct { float delta; int idx; } cand;
static int cand_cmp(const void *a, const void *b)
{
	const cand *x = (const cand *)a, *y = (const cand *)b;
	if (x->delta < y->delta) return -1;
	if (y->delta < x->delta) return 1;
	return (x->idx > y->idx) - (x->idx < y->idx);
}

/* rad.cpp:387-404 listCanditate */
int orc_select(const orc_scene *s, int *out, int cap)
{
	cand *C = (cand *)malloc(sizeof(cand) * (size_t)(s->trinum > 0 ? s->trinum : 1));
	int m = 0;
	for (int i = 0; i < s->trinum; i++) {
		const orc_triangle *tp = &s->tri[i];
		if (tp->parent < 0)
			continue;
		float delta = tp->deltaB[0] + tp->deltaB[1] + tp->deltaB[2];
		if (delta > s->converge) {
			C[m].delta = delta;
			C[m].idx = i;
			m++;
		}
	}
	int n = (m + 4) / 5;
	if (s->parallel_src < n) n = s->parallel_src;
	if (n > 1)
		qsort(C, (size_t)m, sizeof(cand), cand_cmp);
	else if (m > 0) {	/* only the maximum is needed */
		int best = 0;
		for (int i = 1; i < m; i++)
			if (cand_cmp(&C[i], &C[best]) > 0) best = i;
		cand t = C[best]; C[best] = C[m - 1]; C[m - 1] = t;
	}
	int k = 0;
	for (int i = 0, j = m - 1; i < n && k < cap; i++, j--)
		out[k++] = C[j].idx;
	free(C);
	return k;
}

static void stats_add(orc_stats *a, const orc_stats *b)
{
	long long *x = (long long *)a;
	const long long *y = (const long long *)b;
	for (size_t i = 0; i < sizeof(orc_stats) / sizeof(long long); i++) x[i] += y[i];
}

static int is_shooter(const int *C, int n, int i)
{
	for (int j = 0; j < n; j++) if (C[j] == i) return 1;
	return 0;
}

/* rad.cpp:419-439: the body of the receiver loop for slot i. Returns Shade's last result. */
static void receive(orc_scene *s, const int *C, int nC, int i, orc_stats *st)
{
	orc_triangle *des = &s->tri[i];
	if (des->parent < 0)
		return;
	if (is_shooter(C, nC, i))
		return;
	int unsat = 0;
	for (int v = 0; v < 3; v++)
		for (int c = 0; c < 3; c++)
			unsat |= (des->accB[v][c] < 255.0);
	if (!unsat)
		return;
	st->receivers++;
	int logdes = des->parent;
	for (int j = 0; j < nC; j++) {
		int physrc = C[j];
		if (physrc == i)
			continue;
		shade(s, physrc, s->tri[physrc].parent, i, logdes, st);
	}
}

/* rad.cpp:410-449 */
void orc_shot_with(orc_scene *s, const int *C, int nC)
{
	int prev = 0;
	for (;;) {
		int n = s->trinum;
		if (!s->split_enabled) {
			int nthreads = s->threads > 0 ? s->threads : 1;
			(void)nthreads;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
			{
				orc_stats local;
				memset(&local, 0, sizeof(local));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
				for (int i = prev; i < n; i++)
					receive(s, C, nC, i, &local);
#ifdef _OPENMP
#pragma omp critical
#endif
				stats_add(&s->stats, &local);
			}
		} else {
			/* splitting mutates the pool: serial, slot order (the reference is only
			 * deterministic at -jobs 1 on this path) */
			for (int i = prev; i < n; i++)
				receive(s, C, nC, i, &s->stats);
		}
		prev = n;
		if (n == s->trinum)
			break;
	}
	for (int j = 0; j < nC; j++)
		s->tri[C[j]].deltaB[0] = s->tri[C[j]].deltaB[1] = s->tri[C[j]].deltaB[2] = 0.0f;
}

int orc_shot(orc_scene *s, int *shooters_out, int cap, int *nshooters)
{
	int stackbuf[64];
	int n_cap = s->parallel_src > 0 ? s->parallel_src : 1;
	int *C = n_cap <= 64 ? stackbuf : (int *)malloc(sizeof(int) * (size_t)n_cap);
	int nC = orc_select(s, C, n_cap);
	if (nshooters) *nshooters = nC;
	if (shooters_out)
		for (int j = 0; j < nC && j < cap; j++) shooters_out[j] = C[j];
	if (nC > 0)
		orc_shot_with(s, C, nC);
	if (C != stackbuf) free(C);
	return nC > 0;
}

/* rad.cpp:167-178 */
void orc_snapshot_color(const orc_triangle *t, int rgb[3])
{
	for (int j = 0; j < 3; j++) {
		float f = 0;
		for (int k = 0; k < 3; k++)
			f += t->accB[k][j];
		f /= 3;
		float m = (255.0f < f) ? 255.0f : f;	/* std::min(f, 255.0f) */
		f = (m < 0.0f) ? 0.0f : m;		/* std::max(min, 0.0f) */
		if (isnan(f))
			f = 0;
		rgb[j] = (int)f;
	}
}
