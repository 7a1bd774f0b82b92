/*
 * The post-order tree program shared by the fused 4-state kernels (fused4.cuh) and the FP64
 * tensor-pipe kernels (dmma.cu).  Built on the host by build_program (plf_engine.cu) from the
 * csr_graph of the reference (csr_graph.c:47-151) with Sethi-Ullman child ordering.
 */
#pragma once

#define F4_MAXD 3          /* max out-degree handled by the fused kernel */
#define F4_KIND_TIP 0
#define F4_KIND_CUR 1
#define F4_KIND_STACK 2

struct F4Op {
    int node;
    int first_child;
    int nchild;
    int slot;            /* scratch slot of this node's inside vector */
    int code_row;        /* row of the codes tile if the node carries data, else -1 */
    int spill_before;    /* the register-resident partial is not consumed by this op */
};

struct F4Child {
    int kind;
    int slot;            /* scratch slot if internal, else -1 */
    int mat;             /* internal child: index into the compact internal-edge matrices;
                            tip child: index into the compact tip tables */
    int code_row;        /* tip child: row of the codes tile */
    int edge;            /* csr idx (output position) */
    int node;            /* the child node (output position of its marginal) */
    int pad[2];
};

