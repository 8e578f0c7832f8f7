/* A plain-C client of include/mhaq_fq.h: proves the boundary is a C ABI (no C++ / torch types)
 * and exercises the entry points that need no GPU: version, geometry, argument validation.
 * Built and run by tests/test_abi.py::test_plain_c_client_links_and_runs. */
#include <stdio.h>
#include <stddef.h>
#include "mhaq_fq.h"

#define CHECK(cond)                                                   \
    do {                                                              \
        if (!(cond)) {                                                \
            fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #cond); \
            return 1;                                                 \
        }                                                             \
    } while (0)

int main(void) {
    float dummy[4] = {0};
    double wsd[8] = {0};
    unsigned int tk[1] = {0};
    CHECK(mhaq_fq_abi_version() == MHAQ_FQ_ABI_VERSION);
    CHECK(mhaq_fq_build_info() != NULL);
    CHECK(mhaq_fq_num_tasks(1, 4097) == 2);
    CHECK(mhaq_fq_workspace_bytes(4, 1000) >= 4 * 8 * (int64_t)sizeof(double));
    CHECK(mhaq_fq_ticket_count(64, 576, 64) == 64 + 4 + 2048 * 12);
    /* argument errors are reported before anything touches CUDA */
    CHECK(mhaq_fq_fwd_f32(NULL, dummy, NULL, dummy, dummy, NULL, NULL, 0, 0, 0, 0,
                          MHAQ_FQ_PARAMS_LINEAR, 1, 4, 1, NULL, NULL) == MHAQ_FQ_ENULL);
    CHECK(mhaq_fq_fwd_f32(dummy, dummy, NULL, dummy, dummy, NULL, NULL, 3, 0, 0, 0,
                          MHAQ_FQ_PARAMS_LINEAR, 1, 4, 1, NULL, NULL) == MHAQ_FQ_EINVAL);
    CHECK(mhaq_fq_bwd_f32(dummy, dummy, dummy, dummy, dummy, NULL, NULL, 0, 0, 0, 0,
                          MHAQ_FQ_PARAMS_LINEAR, 1, 4, 1, 42, 0, NULL, 0, 0, NULL, NULL, wsd,
                          NULL) == MHAQ_FQ_EINVAL);
    CHECK(mhaq_fq_bwd_finalize_f32(wsd, NULL, NULL, NULL, NULL, NULL, 0, 0, 0, 0,
                                   MHAQ_FQ_PARAMS_LINEAR, 1, 4, 1, dummy, NULL, NULL, NULL,
                                   NULL) == MHAQ_FQ_ENULL);
    CHECK(mhaq_fq_wrow_fwd_f32(NULL, dummy, dummy, 1, 4, NULL, NULL, NULL, NULL) == MHAQ_FQ_ENULL);
    CHECK(mhaq_fq_wrow_fwd_f32(dummy, dummy, dummy, 1, 0, NULL, NULL, NULL, NULL) == MHAQ_FQ_EINVAL);
    CHECK(mhaq_fq_wrow_bwd_f32(dummy, dummy, dummy, dummy, dummy, NULL, NULL, NULL, 1, 4,
                               MHAQ_FQ_AEWGS, NULL, 0, 0, NULL, dummy, dummy, NULL) == MHAQ_FQ_EINVAL);
    CHECK(mhaq_fq_wrow_bwd_f32(dummy, dummy, dummy, NULL, dummy, NULL, NULL, NULL, 1, 4,
                               MHAQ_FQ_STE, NULL, 0, 0, NULL, dummy, dummy, NULL) == MHAQ_FQ_ENULL);
    CHECK(mhaq_fq_rowstat_f32(NULL, 1, 4, NULL, NULL, NULL, NULL, NULL) == MHAQ_FQ_ENULL);
    /* empty tensors are a no-op, not an error */
    CHECK(mhaq_fq_fwd_f32(dummy, dummy, NULL, dummy, dummy, NULL, NULL, 0, 0, 0, 0,
                          MHAQ_FQ_PARAMS_LINEAR, 0, 4, 1, NULL, NULL) == 0);
    CHECK(mhaq_fq_wrow_fwd_f32(dummy, dummy, dummy, 0, 4, NULL, NULL, NULL, NULL) == 0);
    (void)tk;
    printf("abi_client OK: ABI v%d, %s\n", mhaq_fq_abi_version(), mhaq_fq_build_info());
    return 0;
}
