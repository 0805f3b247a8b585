// vmgym_sort.cuh — numpy scalar argsort replay for the best-fit compat tie mode (see DESIGN.md "tie-break").
#pragma once
#include <stdint.h>
namespace vmgym {
// ---- numpy's scalar argsort replayed by one lane (best-fit compat tie mode) ------------
static __device__ __noinline__ void introsort_argsort(const float* v, uint16_t* t, int num)
{
    // third-party algorithm: numpy npysort aquicksort_<float> (median-of-3 quicksort, insertion sort for
    // partitions of <= 16, larger side pushed); the heapsort fallback (depth limit) is kept for completeness.
    int pl = 0, pr = num - 1;
    int st_l[64], st_r[64], st_d[64];
    int sp = 0, cdepth = 0;
    for (int i = 0; i < num; i++) t[i] = (uint16_t)i;
    for (int n = num; n >>= 1;) cdepth++;
    cdepth *= 2;
    for (;;) {
        bool heap = false;
        if (cdepth < 0) heap = true;
        if (!heap) {
            while (pr - pl > 15) {
                int pm = pl + ((pr - pl) >> 1);
                uint16_t x;
                if (v[t[pm]] < v[t[pl]]) { x = t[pm]; t[pm] = t[pl]; t[pl] = x; }
                if (v[t[pr]] < v[t[pm]]) { x = t[pr]; t[pr] = t[pm]; t[pm] = x; }
                if (v[t[pm]] < v[t[pl]]) { x = t[pm]; t[pm] = t[pl]; t[pl] = x; }
                const float vp = v[t[pm]];
                int pi = pl, pj = pr - 1;
                x = t[pm]; t[pm] = t[pj]; t[pj] = x;
                for (;;) {
                    do { ++pi; } while (v[t[pi]] < vp);
                    do { --pj; } while (vp < v[t[pj]]);
                    if (pi >= pj) break;
                    x = t[pi]; t[pi] = t[pj]; t[pj] = x;
                }
                x = t[pi]; t[pi] = t[pr - 1]; t[pr - 1] = x;
                if (pi - pl < pr - pi) { st_l[sp] = pi + 1; st_r[sp] = pr; pr = pi - 1; }
                else { st_l[sp] = pl; st_r[sp] = pi - 1; pl = pi + 1; }
                st_d[sp] = --cdepth;
                sp++;
                if (cdepth < 0) { heap = true; break; }
            }
        }
        if (heap) {
            // heapsort of t[pl..pr] (1-based sift-down on a = t + pl - 1)
            uint16_t* a = t + pl - 1;
            int n = pr - pl + 1, i, j, l;
            uint16_t tmp;
            for (l = n >> 1; l > 0; --l) {
                tmp = a[l];
                for (i = l, j = l << 1; j <= n;) {
                    if (j < n && v[a[j]] < v[a[j + 1]]) j += 1;
                    if (v[tmp] < v[a[j]]) { a[i] = a[j]; i = j; j += j; } else break;
                }
                a[i] = tmp;
            }
            for (; n > 1;) {
                tmp = a[n]; a[n] = a[1]; n -= 1;
                for (i = 1, j = 2; j <= n;) {
                    if (j < n && v[a[j]] < v[a[j + 1]]) j++;
                    if (v[tmp] < v[a[j]]) { a[i] = a[j]; i = j; j += j; } else break;
                }
                a[i] = tmp;
            }
        } else {
            for (int pi = pl + 1; pi <= pr; ++pi) {
                const uint16_t vi = t[pi];
                const float vp = v[vi];
                int pj = pi;
                while (pj > pl && vp < v[t[pj - 1]]) { t[pj] = t[pj - 1]; pj--; }
                t[pj] = vi;
            }
        }
        if (sp == 0) break;
        sp--;
        pl = st_l[sp]; pr = st_r[sp]; cdepth = st_d[sp];
    }
}

}  // namespace vmgym
