// TEST INFRASTRUCTURE (oracle/_ref harness). Shared declarations of the three harness TUs.
#ifndef HARNESS_API_H
#define HARNESS_API_H

// Raw scene dump: header, then TriStore[0..trinum) as the reference's own 236-byte
// `Triangle` records (include/rad.h:25-39), TreeNodeStore[0..nnodes) as 60-byte
// `TreeNode` records (include/rad.h:41-45), then TriListStore[0..nlist) ints.
struct HarnessSceneHeader {
	char magic[8];		// "RADSCN1\0"
	int n0;			// number of input triangles (occluders live in slots [0,n0))
	int trinum;
	int nnodes, nlist;
	int gridnum;
	float mingridlen;
	float G0[3], G1[3];
	float converge;
	int parallel_src;
	float light;
	int tri_limit;
	float area_limit, sample_area, delta_ff;
	int tri_record_bytes, node_record_bytes;
	int reserved[8];
};

int harness_tree_nodes(void);
int harness_tree_listlen(void);
int harness_grid_num(void);
float harness_min_grid_len(void);
const void *harness_tree_node_ptr(void);
const int *harness_tree_list_ptr(void);

void harness_ray_log_open(const char *path);
void harness_ray_log_close(void);
void harness_ray_counters_reset(void);
void harness_ray_counters(long long *rays, long long *visible);
#endif
