// Table of compiled pcg_kernel variants.  Each (group, nodal) pair is its own translation unit (pcg_inst.cu built
// with -DPCG_GROUP=g -DPCG_NODAL=b) so the heavily unrolled kernels compile in parallel.
#pragma once

namespace tfin {

struct PcgVariant {
    int R, WT, WR, maxT, minB, nodal;
    const void* func;      // host stub of pcg_kernel<R, WT, WR, nodal, maxT, minB>
    const void* func_adj;  // ... with the adjoint solves compiled in (nodal variants only, else nullptr)
};

//        X(R, WT, WR, MAXT, MINB)
// group 0/1: two CTAs per SM, values in shared memory (WT = 4 / 8)
#define PCG_GROUP_0(X) X(5, 4, 0, 320, 2) X(6, 4, 0, 288, 2) X(8, 4, 0, 256, 2)
#define PCG_GROUP_1(X) X(5, 8, 0, 320, 2) X(6, 8, 0, 288, 2) X(8, 8, 0, 256, 2)
// group 2: one CTA per SM, values in registers (WT = 4, 6)
#define PCG_GROUP_2(X) X(3, 4, 4, 576, 1) X(4, 4, 4, 448, 1) X(4, 6, 6, 448, 1)
// group 3: large meshes (n <= 8191), one CTA per SM
#define PCG_GROUP_3(X) X(16, 4, 0, 512, 1) X(16, 8, 0, 512, 1) X(16, 12, 0, 512, 1)
// group 4: wide rows (unstructured meshes), two CTAs per SM
#define PCG_GROUP_4(X) X(6, 12, 0, 288, 2) X(8, 12, 0, 256, 2)
#define PCG_NUM_GROUPS 5

const PcgVariant* pcg_variants(int group, int nodal, int* count);

}  // namespace tfin
