// Stress test of csrc/kem_copy_pool.h (tests/test_copy_pool.py): many parallel copies of
// odd sizes must be exact, and the process must EXIT (a destroyed condition variable with
// parked workers once hung interpreter shutdown).
#include "kem_copy_pool.h"

#include <stdio.h>

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 500;
    const size_t n = 8u << 20;
    std::vector<char> a(n), b(n);
    for (int it = 0; it < iters; ++it) {
        for (size_t i = 0; i < n; i += 4097) a[i] = (char)(it + i);
        const size_t len = n - (size_t)(it % 7) * 1001;
        CopyPool::get().copy(b.data(), a.data(), len);
        if (memcmp(a.data(), b.data(), len)) {
            printf("MISMATCH %d\n", it);
            return 1;
        }
    }
    // fills: zero (memset path) and a non-zero value, odd lengths
    std::vector<double> d((n / 8) + 3, -1.0);
    for (double v : {0.0, 2.5, -0.0}) {
        const size_t len = d.size() - 3;
        CopyPool::get().fill(d.data(), v, len);
        for (size_t i = 0; i < len; ++i)
            if (memcmp(&d[i], &v, sizeof v)) {
                printf("FILL MISMATCH %g at %zu\n", v, i);
                return 1;
            }
        if (d[len] != -1.0) {
            printf("FILL OVERRUN\n");
            return 1;
        }
    }
    CopyPool::get().copy(b.data(), a.data(), 0);
    CopyPool::get().copy(b.data(), a.data(), 17);
    printf("COPY_POOL_OK\n");
    return 0;
}
