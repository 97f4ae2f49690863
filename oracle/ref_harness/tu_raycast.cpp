// TEST INFRASTRUCTURE (oracle/_ref harness). Not part of the product.
// Translation unit 2 of 3: the UNMODIFIED reference octree/ray code
// (/root/reference/src/raycast.cpp) plus accessors to its file-static stores
// (raycast.cpp:9-16), which are not linkable from outside.
#include REF_RAYCAST_CPP
#include "harness_api.h"

int harness_tree_nodes(void) { return TreeNodeStorePtr; }
int harness_tree_listlen(void) { return TriListStorePtr; }
int harness_grid_num(void) { return GridNum; }
float harness_min_grid_len(void) { return MinGridLen; }
const void *harness_tree_node_ptr(void) { return TreeNodeStore; }
const int *harness_tree_list_ptr(void) { return TriListStore; }
