// twotime.cu -- two-time correlation for one dynamic bin (reference corr.cpp:781-924).
#include "internal.h"

namespace xpcs {

int launch_twotime(xpcs_handle_s *h, int qbin, int wsize, int method, int average, float *C,
                   float *g2full, float *g2partials, float *sg)
{
    (void)qbin; (void)wsize; (void)method; (void)average; (void)C; (void)g2full; (void)g2partials; (void)sg;
    return fail(h, XPCS_E_STATE, "two-time correlation is not built yet");
}

}  // namespace xpcs
