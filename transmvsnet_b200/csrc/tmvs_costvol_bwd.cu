// placeholder, replaced below
#include "tmvs_common.cuh"
extern "C" size_t tmvs_costvol_bwd_workspace_bytes(int, int, int, int, int, int) { return 0; }
extern "C" int tmvs_costvol_bwd(const float *, int64_t, int64_t, int64_t, int64_t, const float *, const float *,
                                const float *, int, const float *, float *, float *, void *, size_t, int, int, int,
                                int, int, int, tmvs_stream_t) { return TMVS_E_UNSUPPORTED; }
