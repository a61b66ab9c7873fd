// Test harness (CPU only): exposes swift-game-engine_b200/csrc/cq_reftree.h — the product's host rebuild of the
// reference BVH's visiting order — to pytest.
#include "../../swift-game-engine_b200/csrc/cq_reftree.h"

extern "C" {

// lo/hi: n*3 floats.  order_out / rank_out: n ints.  counts_out: [n_nodes, n_leaves].
// nodes_out (may be NULL): 5 ints per node (left, right, start, count, parent), at most max_nodes nodes.
void reftree_build(const float *lo, const float *hi, int n, int threads, int *order_out, int *rank_out, int *counts_out,
                   int *nodes_out, int max_nodes) {
    cq::RefTree T;
    cq::build_ref_tree(lo, hi, 3, n, T, threads);
    for (int i = 0; i < n; i++) order_out[i] = T.order[i], rank_out[i] = T.rank[i];
    counts_out[0] = (int)T.nodes.size(), counts_out[1] = T.nLeaves;
    if (nodes_out)
        for (int k = 0; k < (int)T.nodes.size() && k < max_nodes; k++) {
            const cq::RefNode &nd = T.nodes[k];
            int *o = nodes_out + 5 * k;
            o[0] = nd.left, o[1] = nd.right, o[2] = nd.start, o[3] = nd.count, o[4] = nd.parent;
        }
}
}
