// ref_pileup.cpp -- TEST INFRASTRUCTURE ONLY.  Pins row N3: exposes the reference's OWN threshold rule,
// s_resolve_scaled_prob_threshold (src/app/hifimeth/pileup.cpp:355-436), by including pileup.cpp as it lies under
// /root/reference into this translation unit (the function is static) and calling it through one extern "C" entry.
// Nothing else of pileup.cpp is ever called: its unresolved htslib / FASTA symbols stay lazily bound (oracle/Makefile links
// this object into _ref/libhifimeth_ref_pileup.so without -z now / --no-undefined).
#include <cstddef>
#include <cstdint>
#include <limits>

#include "app/hifimeth/pileup.cpp"

extern "C" void hmref_pileup_thresholds(const uint64_t* cpg, const uint64_t* chg, const uint64_t* chh, uint8_t out[3])
{
    size_t a[256], b[256], c[256];
    for (int i = 0; i < 256; ++i) { a[i] = (size_t)cpg[i]; b[i] = (size_t)chg[i]; c[i] = (size_t)chh[i]; }
    u8 t0 = 0, t1 = 0, t2 = 0;
    ns_pileup::s_resolve_scaled_prob_threshold(a, b, c, t0, t1, t2);
    out[0] = t0; out[1] = t1; out[2] = t2;
}
