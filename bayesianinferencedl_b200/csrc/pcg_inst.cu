// One translation unit per (PCG_GROUP, PCG_NODAL): instantiates its pcg_kernel variants and exports their table.
#include "pcg_small.cuh"
#include "pcg_variants.h"

#ifndef PCG_GROUP
#error "compile with -DPCG_GROUP=<g> -DPCG_NODAL=<0|1>"
#endif
#define PCG_CAT2(a, b) a##b
#define PCG_CAT(a, b) PCG_CAT2(a, b)
#define PCG_LIST PCG_CAT(PCG_GROUP_, PCG_GROUP)

namespace tfin {

#if PCG_NODAL
#define PCG_ADJ_PTR(R, WT, WR, MAXT, MINB) (const void*)&pcg_kernel<R, WT, WR, true, MAXT, MINB, true>
#else
#define PCG_ADJ_PTR(R, WT, WR, MAXT, MINB) nullptr
#endif
#define PCG_ROW(R, WT, WR, MAXT, MINB)                                                                      \
    {R, WT, WR, MAXT, MINB, PCG_NODAL, (const void*)&pcg_kernel<R, WT, WR, (PCG_NODAL != 0), MAXT, MINB>, \
     PCG_ADJ_PTR(R, WT, WR, MAXT, MINB)},
static const PcgVariant k_table[] = {PCG_LIST(PCG_ROW)};

const PcgVariant* PCG_CAT(PCG_CAT(pcg_variants_g, PCG_GROUP), PCG_CAT(_n, PCG_NODAL))(int* count) {
    *count = (int)(sizeof(k_table) / sizeof(k_table[0]));
    return k_table;
}

}  // namespace tfin
