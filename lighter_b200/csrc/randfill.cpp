/*
 * randfill.cpp -- n consecutive draws of the process's libc rand() stream, fast.
 *
 * The reference draws one randf() = rand()/RAND_MAX per lumel in the AO pass (lighter.cpp:819,
 * lighter_int.hpp:271) and one per ltr_LightAdd (lighter.cpp:1300); the bake replays that stream on the host
 * so that a caller's srand() is honoured and results equal the reference's.  rand() takes a lock per call
 * (~12 ns): 7.8 M lumels = ~0.1 s of host time, which an 8-GPU bake cannot hide behind its GPU stages.
 *
 * glibc's rand() is the TYPE_3 additive-feedback generator of random(3): 31 words of state,
 *   r[f] += r[r_];  result = r[f] >> 1;  f, r_ advance modulo 31  (f = r_ + 3).
 * setstate(3) hands out the live state table (it stores the rear index and the type in the word before the
 * table and returns a pointer to it), so the stream can be advanced on the table directly and handed back:
 * the libc state after rand_fill(n) is exactly the state after n calls of rand().  Everything is checked
 * against the real rand() once per process; on any mismatch (another libc) the plain loop is used.
 */
#include "randfill.h"

#include <stdlib.h>
#include <string.h>

#include <mutex>

namespace {

const int DEG = 31, SEP = 3, MAX_TYPES = 5, TYPE_3 = 3;

/* advance the generator in `info` (info[0] = 5*rear + type, info[1..31] = table) by n draws */
bool advance(int32_t *info, float *out, uint64_t n)
{
    if (info[0] % MAX_TYPES != TYPE_3) return false;
    int r = info[0] / MAX_TYPES;
    if (r < 0 || r >= DEG) return false;
    uint32_t *tbl = reinterpret_cast<uint32_t *>(info + 1);
    int f = (r + SEP) % DEG;
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t val = tbl[f] += tbl[r];
        out[i] = (float)(int)(val >> 1) / (float)RAND_MAX;
        if (++f == DEG) f = 0;
        if (++r == DEG) r = 0;
    }
    info[0] = MAX_TYPES * r + TYPE_3;
    return true;
}

alignas(8) char g_scratch[128];

/* fast path on the live libc state; false = not applicable (state untouched) */
bool fill_fast(float *out, uint64_t n)
{
    char *old = initstate(1u, g_scratch, sizeof(g_scratch));     /* libc now runs on the scratch state; `old` = the live one */
    if (!old) return false;
    const bool ok = advance(reinterpret_cast<int32_t *>(old), out, n);
    setstate(old);                                               /* hand the (advanced) state back */
    return ok;
}

bool self_check()
{
#if defined(__GLIBC__)
    alignas(8) static char a[128], b[128];
    float ea[96], eb[96];
    char *old = initstate(20261017u, a, sizeof(a));              /* reference stream: real rand() on state a */
    if (!old) return false;
    for (int i = 0; i < 64; ++i) ea[i] = (float)rand() / (float)RAND_MAX;
    initstate(20261017u, b, sizeof(b));                          /* same seed on state b, advanced by hand ... */
    char *live_b = initstate(1u, g_scratch, sizeof(g_scratch));  /* (switch away so that b is not live while we edit it) */
    bool ok = live_b == b && advance(reinterpret_cast<int32_t *>(live_b), eb, 40);
    setstate(b);
    for (int i = 40; i < 64 && ok; ++i) eb[i] = (float)rand() / (float)RAND_MAX;    /* ... then continued by the real rand() */
    setstate(old);
    return ok && memcmp(ea, eb, 64 * sizeof(float)) == 0;
#else
    return false;
#endif
}

} // namespace

bool rand_fill(float *out, uint64_t n)
{
    static std::once_flag once;
    static bool fast = false;
    std::call_once(once, []() { fast = self_check() && !getenv("LTR_RAND_SLOW"); });
    if (fast && fill_fast(out, n)) return true;
    for (uint64_t i = 0; i < n; ++i) out[i] = (float)rand() / (float)RAND_MAX;
    return false;
}
