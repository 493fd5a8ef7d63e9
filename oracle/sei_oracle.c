/*
 * sei_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See sei_oracle_impl.h.
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC) -> oracle/libsei_oracle.so
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define FN(name) CAT(name, _f32)
#include "sei_oracle_impl.h"
#undef REAL
#undef FN

#define REAL double
#define FN(name) CAT(name, _f64)
#include "sei_oracle_impl.h"
#undef REAL
#undef FN

/* get_kernel (src/physics/kernels.py:3-28): named blur kernels in float64.
 * Gaussian_R{1,2,3}: size 6R+1, exp(-(u^2+v^2)/(2R^2)) normalised to sum 1.
 * Box_R{2,3,4}: size 2R+1, constant 1/size^2.  Returns the size, or -1 if unknown. */
int orc_named_kernel(const char* name, double* out)
{
    int level = 0;
    if (sscanf(name, "Gaussian_R%d", &level) == 1 && level >= 1 && level <= 3 && strlen(name) == 11) {
        const int k = 6 * level + 1;
        double sum = 0;
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                const double u = i - (k - 1) / 2.0, v = j - (k - 1) / 2.0;
                out[i * k + j] = exp(-(u * u + v * v) / (2.0 * level * level));
                sum += out[i * k + j];
            }
        for (int i = 0; i < k * k; ++i) out[i] /= sum;
        return k;
    }
    if (sscanf(name, "Box_R%d", &level) == 1 && level >= 2 && level <= 4 && strlen(name) == 6) {
        const int k = 2 * level + 1;
        for (int i = 0; i < k * k; ++i) out[i] = 1.0 / (double)(k * k);
        return k;
    }
    return -1;
}

int orc_abi_version(void) { return 1; }

/* torchrun exports OMP_NUM_THREADS=1; the CPU-baseline leg of bench.py asks for every host core explicitly */
#ifdef _OPENMP
#include <omp.h>
int orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
int orc_set_threads(int n) { (void)n; return 1; }
#endif
