// C ABI of the host front end (include/radk.h, "host front end" section).
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include "model.hpp"

struct radk_model { radk::Model m; };

namespace radk {
static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
}  // namespace radk

extern "C" {

const char *radk_last_error(void) { return radk::g_error.c_str(); }
int radk_version(void) { return RADK_VERSION; }

void radk_default_opts(radk_opts *o) {
	memset(o, 0, sizeof(*o));
	o->converge = 1200.0f;		// rad.cpp:20
	o->parallel_src = 1;		// rad.cpp:30
	o->tri_limit = 1000000;		// rad.cpp:22
	o->area_limit = 5.0f;		// rad.cpp:18
	o->sample_area = -1.0f;		// rad.cpp:19
	o->delta_ff = 0.000005f;	// rad.cpp:21
	o->max_tri = radk::kRefMaxTri;
	o->world = 1;
}

void radk_default_model_opts(radk_model_opts *o) {
	memset(o, 0, sizeof(*o));
	o->light_scale = 1.0f;		// rad.cpp:25
	o->tri_limit = 1000000;
	o->max_tri = radk::kRefMaxTri;
	o->max_tree_nodes = radk::kRefMaxTreeNode;
	o->prepartition = 1;		// config.h:8 PRE_PARTITION
	o->sample_area = -1.0f;
}

int radk_model_parse(const char *text, size_t len, const radk_model_opts *o, radk_model **out) {
	if (!text || !out) { radk::set_error("radk_model_parse: null argument"); return RADK_ERR_ARG; }
	radk_model *h = new (std::nothrow) radk_model();
	if (!h) { radk::set_error("out of memory"); return RADK_ERR_LIMIT; }
	radk::Model &m = h->m;
	if (o) m.opts = *o; else radk_default_model_opts(&m.opts);
	int rc = radk::read_triangles(text, len, m);
	if (rc == RADK_OK) rc = radk::build_octree(m);		// InitModelFromFile, rad.cpp:244-252
	if (rc != RADK_OK) { delete h; return rc; }
	radk::init_radiosity(m);				// DoRadiosity, rad.cpp:493
	if (m.opts.prepartition) radk::prepartition(m);		// rad.cpp:494-496
	m.tri.resize((size_t)m.trinum);
	m.tri.shrink_to_fit();
	*out = h;
	return RADK_OK;
}

int radk_model_load(const char *tri_path, const radk_model_opts *o, radk_model **out) {
	if (!tri_path || !out) { radk::set_error("radk_model_load: null argument"); return RADK_ERR_ARG; }
	FILE *f = fopen(tri_path, "rb");
	if (!f) { radk::set_error(std::string("File open error::") + tri_path); return RADK_ERR_IO; }
	std::string text;
	char buf[1 << 16];
	size_t n;
	while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
	fclose(f);
	return radk_model_parse(text.data(), text.size(), o, out);
}

void radk_model_free(radk_model *m) {
	if (m && m->m.writer.joinable()) m->m.writer.join();
	delete m;
}

int radk_model_scene(const radk_model *h, radk_scene *out) {
	if (!h || !out) { radk::set_error("radk_model_scene: null argument"); return RADK_ERR_ARG; }
	const radk::Model &m = h->m;
	out->tri = m.tri.data();
	out->n0 = m.n0;
	out->trinum = m.trinum;
	out->nodes = m.nodes.data();
	out->nnodes = (int)m.nodes.size();
	out->list = m.list.data();
	out->nlist = (int)m.list.size();
	for (int k = 0; k < 3; k++) out->G0[k] = m.G0[k], out->G1[k] = m.G1[k];
	out->min_grid_len = m.min_grid_len;
	out->grid_num = m.grid_num;
	return RADK_OK;
}

float radk_model_sample_area(const radk_model *h) { return h ? h->m.sample_area : -1.0f; }
float radk_model_min_area(const radk_model *h) { return h ? h->m.min_area : 0.0f; }
float radk_model_max_area(const radk_model *h) { return h ? h->m.max_area : 0.0f; }

radk_triangle *radk_model_triangles(radk_model *h, int *trinum) {
	if (!h) return nullptr;
	if (trinum) *trinum = h->m.trinum;
	return h->m.tri.data();
}

int radk_model_set_triangles(radk_model *h, const radk_triangle *tri, int trinum) {
	if (!h || !tri || trinum < 0) { radk::set_error("radk_model_set_triangles: bad argument"); return RADK_ERR_ARG; }
	if (h->m.writer.joinable()) h->m.writer.join();
	h->m.tri.assign(tri, tri + trinum);
	h->m.trinum = trinum;
	h->m.geom_valid = false;
	return RADK_OK;
}

int radk_model_dump(const radk_model *h, const radk_opts *o, const char *path) {
	if (!h || !path) { radk::set_error("radk_model_dump: null argument"); return RADK_ERR_ARG; }
	const radk::Model &m = h->m;
	struct Header {		// oracle/ref_harness/harness_api.h "RADSCN1"
		char magic[8];
		int n0, trinum, nnodes, nlist, gridnum;
		float mingridlen, G0[3], G1[3], converge;
		int parallel_src;
		float light;
		int tri_limit;
		float area_limit, sample_area, delta_ff;
		int tri_record_bytes, node_record_bytes, reserved[8];
	} hd;
	memset(&hd, 0, sizeof(hd));
	memcpy(hd.magic, "RADSCN1", 8);
	radk_opts defaults;
	radk_default_opts(&defaults);
	if (!o) o = &defaults;
	hd.n0 = m.n0, hd.trinum = m.trinum, hd.nnodes = (int)m.nodes.size(), hd.nlist = (int)m.list.size();
	hd.gridnum = m.grid_num, hd.mingridlen = m.min_grid_len;
	for (int k = 0; k < 3; k++) hd.G0[k] = m.G0[k], hd.G1[k] = m.G1[k];
	hd.converge = o->converge, hd.parallel_src = o->parallel_src, hd.light = m.opts.light_scale;
	hd.tri_limit = m.opts.tri_limit, hd.area_limit = o->area_limit;
	hd.sample_area = m.sample_area, hd.delta_ff = o->delta_ff;
	hd.tri_record_bytes = (int)sizeof(radk_triangle), hd.node_record_bytes = (int)sizeof(radk_tree_node);
	FILE *f = fopen(path, "wb");
	if (!f) { radk::set_error(std::string("cannot open ") + path); return RADK_ERR_IO; }
	fwrite(&hd, sizeof(hd), 1, f);
	fwrite(m.tri.data(), sizeof(radk_triangle), (size_t)m.trinum, f);
	fwrite(m.nodes.data(), sizeof(radk_tree_node), m.nodes.size(), f);
	fwrite(m.list.data(), sizeof(int), m.list.size(), f);
	fclose(f);
	return RADK_OK;
}

int radk_model_write_wait(radk_model *h) {
	if (!h) { radk::set_error("radk_model_write_wait: null argument"); return RADK_ERR_ARG; }
	if (h->m.writer.joinable()) h->m.writer.join();
	const int rc = h->m.writer_rc;
	h->m.writer_rc = 0;
	if (rc) radk::set_error("output file open failed");
	return rc;
}

int radk_model_write_colors(radk_model *h, const char *prefix, int loop, const int *rgb, int trinum) {
	if (!h || !prefix || !rgb) { radk::set_error("radk_model_write_colors: null argument"); return RADK_ERR_ARG; }
	radk::Model &m = h->m;
	if (trinum != m.trinum) { radk::set_error("radk_model_write_colors: colour array does not match the model"); return RADK_ERR_ARG; }
	const int prev = radk_model_write_wait(h);
	if (!m.geom_valid) radk::build_geometry_cache(m);
	std::string fn(prefix);
	if (loop == 0) fn += ".fin";
	else fn += "." + std::to_string(loop);
	std::string text;
	radk::format_fin_colors(m, rgb, text);
	m.writer = std::thread([&m, fn, loop, text = std::move(text)]() {
		FILE *f = fopen(fn.c_str(), "w");
		if (!f) {	// rad.cpp:131-138
			fprintf(stderr, "Error: output file %s open failed\n", fn.c_str());
			if (loop != 0) { m.writer_rc = RADK_ERR_IO; return; }
			fwrite(text.data(), 1, text.size(), stdout);
			return;
		}
		fwrite(text.data(), 1, text.size(), f);
		fclose(f);
	});
	return prev;
}

int radk_model_format(const radk_model *h, char *buf, size_t cap, size_t *len) {
	if (!h || !len) { radk::set_error("radk_model_format: null argument"); return RADK_ERR_ARG; }
	std::string text;
	radk::format_fin(h->m, text);
	*len = text.size();
	if (buf && cap >= text.size()) memcpy(buf, text.data(), text.size());
	return RADK_OK;
}

int radk_model_write(const radk_model *h, const char *prefix, int loop) {
	if (!h || !prefix) { radk::set_error("radk_model_write: null argument"); return RADK_ERR_ARG; }
	std::string fn(prefix);			// rad.cpp:123-129
	if (loop == 0) fn += ".fin";
	else fn += "." + std::to_string(loop);
	std::string text;
	radk::format_fin(h->m, text);
	FILE *f = fopen(fn.c_str(), "w");
	if (!f) {				// rad.cpp:131-138
		fprintf(stderr, "Error: output file %s open failed\n", fn.c_str());
		if (loop != 0) { radk::set_error("output file open failed"); return RADK_ERR_IO; }
		fwrite(text.data(), 1, text.size(), stdout);
		return RADK_OK;
	}
	fwrite(text.data(), 1, text.size(), f);
	fclose(f);
	return RADK_OK;
}

}  // extern "C"
