/*
 * oracle/dtw_oracle.c -- CPU restatement of the DTW the reference calls.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under abnet3_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * What it restates
 * ----------------
 * The reference does not contain its DTW: abnet3/utils.py:14 imports
 * `DTW` from the third-party Cython module DTW_Cython
 * (requirements.txt:9, `git+https://github.com/Rachine/DTW_Cython.git`, NO
 * commit pin) and calls it exactly once, abnet3/utils.py:149-151:
 *
 *     _, _, paths = DTW(feat1, feat2, return_alignment=True,
 *                       dist_array=distance_array)
 *     path1, path2 = paths[1:]
 *
 * That module is absent from /root/reference and cannot be fetched here, so
 * this file restates the published algorithm of that lineage (ABXpy
 * `dtw.pyx` / abnet `dtw.pyx`): classic symmetric DTW on a precomputed
 * float64 local-distance matrix,
 *
 *     C[0,0] = D[0,0]
 *     C[i,0] = D[i,0] + C[i-1,0]          C[0,j] = D[0,j] + C[0,j-1]
 *     C[i,j] = D[i,j] + min(C[i-1,j], C[i-1,j-1], C[i,j-1])
 *
 * followed by a traceback from (n1-1, n2-1) to (0,0) that, at every interior
 * cell, moves to argmin(C[i-1,j-1], C[i-1,j], C[i,j-1]) taking the FIRST
 * minimum in that order (diagonal, then i-1 "up", then j-1 "left") -- the
 * numpy `argmin` convention of the `_traceback` helper in that lineage.
 *
 * PARITY UNPINNED: the reference holds no golden vector, known-answer test or
 * fixture for DTW costs or paths, and the module's
 * source cannot be inspected here, so the tie rule above is DEFINED by this
 * repository.  On inputs without exact cost ties every symmetric-DTW
 * implementation returns the same path; exact ties are counted separately in
 * the parity tests.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* Direction codes shared with the CUDA kernels (abnet3_b200/csrc). */
enum { ABN_DIR_DIAG = 0, ABN_DIR_UP = 1, ABN_DIR_LEFT = 2 };

/*
 * D       : n1 x n2 local distances, row-major with leading dimension ldd
 * path1/2 : out, capacity >= n1 + n2 - 1, filled in forward order (0,0)->end
 * cost    : out, C[n1-1, n2-1]
 * acc     : optional out, n1 x n2 accumulated-cost matrix (NULL to skip)
 * ties    : optional out, number of interior traceback steps where the
 *           minimum of the three predecessors was not unique
 * returns : path length L (max(n1,n2) <= L <= n1+n2-1), or -1 when D holds a
 *           NaN or a negative entry (the reference asserts d >= 0 before DTW,
 *           abnet3/utils.py:59, and its callers drop the pair,
 *           abnet3/dataloader.py:188-191), or -2 on allocation failure.
 */
int abn_oracle_dtw(const double *D, int n1, int n2, int ldd,
                   int32_t *path1, int32_t *path2, double *cost,
                   double *acc, int32_t *ties)
{
    if (n1 <= 0 || n2 <= 0) return -1;
    for (int i = 0; i < n1; ++i)
        for (int j = 0; j < n2; ++j)
            if (!(D[(size_t)i * ldd + j] >= 0.0)) return -1;

    double *C = acc ? acc : (double *)malloc(sizeof(double) * (size_t)n1 * n2);
    if (!C) return -2;
#define CC(i, j) C[(size_t)(i) * n2 + (j)]
#define DD(i, j) D[(size_t)(i) * ldd + (j)]
    CC(0, 0) = DD(0, 0);
    for (int i = 1; i < n1; ++i) CC(i, 0) = DD(i, 0) + CC(i - 1, 0);
    for (int j = 1; j < n2; ++j) CC(0, j) = DD(0, j) + CC(0, j - 1);
    for (int i = 1; i < n1; ++i) {
        for (int j = 1; j < n2; ++j) {
            double up = CC(i - 1, j), dg = CC(i - 1, j - 1), lf = CC(i, j - 1);
            double m = dg;
            if (up < m) m = up;
            if (lf < m) m = lf;
            CC(i, j) = DD(i, j) + m;
        }
    }
    if (cost) *cost = CC(n1 - 1, n2 - 1);

    /* traceback, written backwards then reversed in place */
    int i = n1 - 1, j = n2 - 1, L = 0, nties = 0;
    path1[L] = i; path2[L] = j; ++L;
    while (i > 0 || j > 0) {
        if (i == 0) {
            --j;
        } else if (j == 0) {
            --i;
        } else {
            double dg = CC(i - 1, j - 1), up = CC(i - 1, j), lf = CC(i, j - 1);
            int dir;
            if (dg <= up && dg <= lf) dir = ABN_DIR_DIAG;
            else if (up <= lf)        dir = ABN_DIR_UP;
            else                      dir = ABN_DIR_LEFT;
            double m = dg < up ? dg : up; if (lf < m) m = lf;
            if ((dg == m) + (up == m) + (lf == m) > 1) ++nties;
            if (dir == ABN_DIR_DIAG) { --i; --j; }
            else if (dir == ABN_DIR_UP) --i;
            else --j;
        }
        path1[L] = i; path2[L] = j; ++L;
    }
    for (int a = 0, b = L - 1; a < b; ++a, --b) {
        int32_t t = path1[a]; path1[a] = path1[b]; path1[b] = t;
        t = path2[a]; path2[a] = path2[b]; path2[b] = t;
    }
    if (ties) *ties = nties;
#undef CC
#undef DD
    if (!acc) free(C);
    return L;
}

/*
 * Cost of a GIVEN monotone path under D (sum of D over the path cells): used
 * by the end-to-end parity tests to re-score a GPU path under the oracle's
 * distance matrix.  Returns NaN if the path is not
 * a valid DTW path from (0,0) to (n1-1,n2-1).
 */
double abn_oracle_path_cost(const double *D, int n1, int n2, int ldd,
                            const int32_t *path1, const int32_t *path2, int L)
{
    if (L <= 0 || path1[0] != 0 || path2[0] != 0 ||
        path1[L - 1] != n1 - 1 || path2[L - 1] != n2 - 1)
        return NAN;
    double s = 0.0;
    for (int k = 0; k < L; ++k) {
        if (k > 0) {
            int di = path1[k] - path1[k - 1], dj = path2[k] - path2[k - 1];
            if (di < 0 || dj < 0 || di > 1 || dj > 1 || di + dj == 0) return NAN;
        }
        s += D[(size_t)path1[k] * ldd + path2[k]];
    }
    return s;
}
