// dsp_fast.cu — time-tiled ("FAST" mode) decimator.  Placeholder until the tiled kernel lands.
#include "common.h"
#include "kernels.h"
namespace tb200 {
int launch_decimate_fast(tb200_handle h, const void *x, int n_chunks, long n, long chunk_stride, int q,
                         const DecimDesign &d, void *dec_out, long n_dec, cudaStream_t st)
{
    (void)h; (void)x; (void)n_chunks; (void)n; (void)chunk_stride; (void)q; (void)d; (void)dec_out; (void)n_dec; (void)st;
    return fail(TB200_E_ARG, "TB200_MODE_FAST is not available in this build");
}
}  // namespace tb200
