/*
 * marlin_oracle.c — CPU restatement of the arithmetic on PasaLab/marlin's dense multiply path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under marlin_b200/ may import, link or call this file; it is
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as
 * the checker and the reported CPU baseline.
 *
 * What it restates (paths relative to the reference's src/main/scala/edu/nju/pasalab/marlin/):
 *   - SubMatrix.multiply (matrix/SubMatrix.scala:87-91) = Breeze `DenseMatrix * DenseMatrix`
 *     = netlib-java BLAS.dgemm(transA, transB, m, n, k, 1.0, a, aOff, lda, b, bOff, ldb, 0.0, c, 0, m).
 *     Breeze 0.11.2 (pom.xml:42) and netlib-java (core 1.1.2, pulled in transitively by Breeze) are
 *     third-party and NOT in /root/reference.  netlib-java's default backend is F2J, a mechanical
 *     Java translation of reference-BLAS dgemm.f (README.md:31 says the pure-Java path is what runs
 *     unless native BLAS is installed).  mo_dgemm_f2j below follows dgemm.f's loop nests and —
 *     because the JVM never contracts a*b+c into an FMA — is compiled with -ffp-contract=off so every
 *     multiply and add rounds separately, exactly as F2J does.
 *   - SubMatrix.multiply(v: Vector) (matrix/SubMatrix.scala:131-139) = Breeze `BDM * BDV` = netlib dgemv, and
 *     DistributedVector.multiply's `v.t * w` = ddot (matrix/DistributedVector.scala:167): mo_dgemv_f2j, mo_ddot_f2j.
 *   - SubMatrix.add / subtract / scalar ops (matrix/SubMatrix.scala:41-85,123-131) = Breeze
 *     element-wise operators; BlockMatrix.transpose's `denseBlock.t.copy` (matrix/BlockMatrix.scala:517).
 *   - MTUtils.hashSeed (utils/MTUtils.scala:18-21), XORShiftRandom.next
 *     (utils/RandomDataGenerator.scala:113-131), java.util.Random.nextLong / nextDouble (JDK),
 *     UniformGenerator.nextValue (utils/RandomDataGenerator.scala:53-65).
 *   - MTUtils.splitMethod / dimToSplit (utils/MTUtils.scala:150-175,204-213).
 *
 * Pinning: the reference cannot run here (no JVM).  The oracle is pinned against every golden vector
 * the reference's own suite holds for this path (DistributedMatrixSuite.scala, exact small-integer
 * results) in tests/test_oracle_golden.py.  For non-integer data and for the random generator the
 * reference has no golden vectors: those parts are "parity unpinned" (see DESIGN.md).
 */
#include <stdint.h>
#include <string.h>

/* reference-BLAS dgemm.f loop order, column-major, no FMA.  Returns 0, or the 1-based index of
 * the bad argument like xerbla. */
int mo_dgemm_f2j(char transa, char transb, int m, int n, int k, double alpha, const double* a, long a_off, int lda,
                 const double* b, long b_off, int ldb, double beta, double* c, long c_off, int ldc) {
    const int nota = (transa == 'N' || transa == 'n');
    const int notb = (transb == 'N' || transb == 'n');
    const int nrowa = nota ? m : k;
    const int nrowb = notb ? k : n;
    if (!nota && !(transa == 'T' || transa == 't' || transa == 'C' || transa == 'c')) return 1;
    if (!notb && !(transb == 'T' || transb == 't' || transb == 'C' || transb == 'c')) return 2;
    if (m < 0) return 3;
    if (n < 0) return 4;
    if (k < 0) return 5;
    if (lda < (nrowa > 1 ? nrowa : 1)) return 8;
    if (ldb < (nrowb > 1 ? nrowb : 1)) return 10;
    if (ldc < (m > 1 ? m : 1)) return 13;
    if (m == 0 || n == 0 || ((alpha == 0.0 || k == 0) && beta == 1.0)) return 0;
    a += a_off;
    b += b_off;
    c += c_off;
#define A(i, j) a[(long)(i) + (long)(j) * lda]
#define B(i, j) b[(long)(i) + (long)(j) * ldb]
#define Cc(i, j) c[(long)(i) + (long)(j) * ldc]
    if (alpha == 0.0) {
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) Cc(i, j) = (beta == 0.0) ? 0.0 : beta * Cc(i, j);
        return 0;
    }
    if (notb) {
        if (nota) { /* C := alpha*A*B + beta*C */
            for (int j = 0; j < n; ++j) {
                if (beta == 0.0) { for (int i = 0; i < m; ++i) Cc(i, j) = 0.0; }
                else if (beta != 1.0) { for (int i = 0; i < m; ++i) Cc(i, j) = beta * Cc(i, j); }
                for (int l = 0; l < k; ++l) {
                    if (B(l, j) != 0.0) {
                        const double temp = alpha * B(l, j);
                        for (int i = 0; i < m; ++i) Cc(i, j) = Cc(i, j) + temp * A(i, l);
                    }
                }
            }
        } else { /* C := alpha*A**T*B + beta*C */
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < m; ++i) {
                    double temp = 0.0;
                    for (int l = 0; l < k; ++l) temp = temp + A(l, i) * B(l, j);
                    Cc(i, j) = (beta == 0.0) ? alpha * temp : alpha * temp + beta * Cc(i, j);
                }
        }
    } else {
        if (nota) { /* C := alpha*A*B**T + beta*C */
            for (int j = 0; j < n; ++j) {
                if (beta == 0.0) { for (int i = 0; i < m; ++i) Cc(i, j) = 0.0; }
                else if (beta != 1.0) { for (int i = 0; i < m; ++i) Cc(i, j) = beta * Cc(i, j); }
                for (int l = 0; l < k; ++l) {
                    if (B(j, l) != 0.0) {
                        const double temp = alpha * B(j, l);
                        for (int i = 0; i < m; ++i) Cc(i, j) = Cc(i, j) + temp * A(i, l);
                    }
                }
            }
        } else { /* C := alpha*A**T*B**T + beta*C */
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < m; ++i) {
                    double temp = 0.0;
                    for (int l = 0; l < k; ++l) temp = temp + A(l, i) * B(j, l);
                    Cc(i, j) = (beta == 0.0) ? alpha * temp : alpha * temp + beta * Cc(i, j);
                }
        }
    }
#undef A
#undef B
#undef Cc
    return 0;
}

/* reference-BLAS dgemv.f (netlib-java F2J), unit strides: y := alpha*op(A)*x + beta*y.  Breeze `BDM * BDV`
 * (SubMatrix.multiply(v: Vector), matrix/SubMatrix.scala:131-139) calls it with alpha = 1, beta = 0. */
int mo_dgemv_f2j(char trans, int m, int n, double alpha, const double* a, long a_off, int lda, const double* x,
                 double beta, double* y) {
    const int nota = (trans == 'N' || trans == 'n');
    if (!nota && !(trans == 'T' || trans == 't' || trans == 'C' || trans == 'c')) return 1;
    if (m < 0) return 2;
    if (n < 0) return 3;
    if (lda < (m > 1 ? m : 1)) return 6;
    if (m == 0 || n == 0 || (alpha == 0.0 && beta == 1.0)) return 0;
    a += a_off;
    const int leny = nota ? m : n;
    if (beta != 1.0) {
        if (beta == 0.0) for (int i = 0; i < leny; ++i) y[i] = 0.0;
        else for (int i = 0; i < leny; ++i) y[i] = beta * y[i];
    }
    if (alpha == 0.0) return 0;
    if (nota) {
        for (int j = 0; j < n; ++j) {
            if (x[j] != 0.0) {
                const double temp = alpha * x[j];
                for (int i = 0; i < m; ++i) y[i] = y[i] + temp * a[(long)i + (long)j * lda];
            }
        }
    } else {
        for (int j = 0; j < n; ++j) {
            double temp = 0.0;
            for (int i = 0; i < m; ++i) temp = temp + a[(long)i + (long)j * lda] * x[i];
            y[j] = y[j] + alpha * temp;
        }
    }
    return 0;
}

/* reference-BLAS ddot.f (unit strides: clean-up loop of n mod 5, then five products per statement, summed left to
 * right).  Breeze `v.t * w` (matrix/DistributedVector.scala:167). */
double mo_ddot_f2j(long n, const double* dx, const double* dy) {
    double dtemp = 0.0;
    if (n <= 0) return dtemp;
    const long m = n % 5;
    for (long i = 0; i < m; ++i) dtemp = dtemp + dx[i] * dy[i];
    if (n < 5) return dtemp;
    for (long i = m; i < n; i += 5)
        dtemp = dtemp + dx[i] * dy[i] + dx[i + 1] * dy[i + 1] + dx[i + 2] * dy[i + 2] + dx[i + 3] * dy[i + 3] +
                dx[i + 4] * dy[i + 4];
    return dtemp;
}

/* Breeze element-wise binary operators on packed arrays (SubMatrix.add/subtract, dotProduct). op: 0 +, 1 -, 2 * */
void mo_binary(int op, long n, const double* a, const double* b, double* out) {
    if (op == 0) for (long i = 0; i < n; ++i) out[i] = a[i] + b[i];
    else if (op == 1) for (long i = 0; i < n; ++i) out[i] = a[i] - b[i];
    else for (long i = 0; i < n; ++i) out[i] = a[i] * b[i];
}

/* denseBlock.t.copy : in is rows x cols column-major (ld), out is cols x rows column-major packed */
void mo_transpose_copy(int rows, int cols, const double* in, long ld, double* out) {
    for (int c = 0; c < cols; ++c)
        for (int r = 0; r < rows; ++r) out[(long)c + (long)r * cols] = in[(long)r + (long)c * ld];
}

/* ---- scala.util.hashing.MurmurHash3.bytesHash(data) (Scala 2.10 library; seed arraySeed) ---- */
static uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static uint32_t mix_last(uint32_t h, uint32_t k) {
    k *= 0xcc9e2d51u; k = rotl32(k, 15); k *= 0x1b873593u;
    return h ^ k;
}
static uint32_t mix(uint32_t h, uint32_t k) {
    h = mix_last(h, k); h = rotl32(h, 13);
    return h * 5u + 0xe6546b64u;
}
static uint32_t avalanche(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
int32_t mo_murmur3_bytes_hash(const uint8_t* data, int len) {
    uint32_t h = 0x3c074a61u;
    int i = 0, rem = len;
    while (rem >= 4) {
        uint32_t k = (uint32_t)data[i] | ((uint32_t)data[i + 1] << 8) | ((uint32_t)data[i + 2] << 16) | ((uint32_t)data[i + 3] << 24);
        h = mix(h, k); i += 4; rem -= 4;
    }
    uint32_t k = 0;
    if (rem == 3) k ^= (uint32_t)data[i + 2] << 16;
    if (rem >= 2) k ^= (uint32_t)data[i + 1] << 8;
    if (rem >= 1) { k ^= data[i]; h = mix_last(h, k); }
    return (int32_t)avalanche(h ^ (uint32_t)len);
}

/* MTUtils.hashSeed (utils/MTUtils.scala:18-21): ByteBuffer.allocate(Long.SIZE = 64 bytes).putLong(seed) */
int64_t mo_hash_seed(int64_t seed) {
    uint8_t buf[64];
    memset(buf, 0, sizeof buf);
    for (int i = 0; i < 8; ++i) buf[i] = (uint8_t)((uint64_t)seed >> (56 - 8 * i));
    return (int64_t)mo_murmur3_bytes_hash(buf, 64);
}

/* XORShiftRandom.next(bits) (utils/RandomDataGenerator.scala:121-127) */
static int32_t xs_next(uint64_t* seed, int bits) {
    uint64_t s = *seed ^ (*seed << 21);
    s ^= (s >> 35);
    s ^= (s << 4);
    *seed = s;
    return (int32_t)(s & ((1ull << bits) - 1));
}

/* UniformGenerator(lo,hi) after setSeed(partition_seed): n successive nextValue() starting with
 * value index `first` of the stream (java.util.Random.nextDouble = ((next(26) << 27) + next(27)) * 2^-53). */
void mo_uniform_fill(int64_t partition_seed, long first, long n, double lo, double hi, double* out) {
    uint64_t s = (uint64_t)mo_hash_seed(partition_seed);
    for (long i = 0; i < first; ++i) { xs_next(&s, 26); xs_next(&s, 27); }
    const double span = hi - lo;
    for (long i = 0; i < n; ++i) {
        const int64_t a = xs_next(&s, 26);
        const int64_t b = xs_next(&s, 27);
        const double u = (double)((a << 27) + b) * 0x1.0p-53;
        out[i] = span * u + lo;
    }
}

/* The same stream continued from an explicit XORShift state (test helper: lets the Python side skip ahead with its own
 * GF(2) matrix power instead of stepping 2*first times). */
void mo_uniform_fill_from_state(uint64_t state, long n, double lo, double hi, double* out) {
    uint64_t s = state;
    const double span = hi - lo;
    for (long i = 0; i < n; ++i) {
        const int64_t a = xs_next(&s, 26);
        const int64_t b = xs_next(&s, 27);
        const double u = (double)((a << 27) + b) * 0x1.0p-53;
        out[i] = span * u + lo;
    }
}

/* java.util.Random(seed): successive nextLong() — the partition seeds of rdd/RandomRDD.scala:28-45 */
void mo_java_random_longs(int64_t seed, int n, int64_t* out) {
    const uint64_t mask = (1ull << 48) - 1;
    uint64_t s = ((uint64_t)seed ^ 0x5DEECE66Dull) & mask;
    for (int i = 0; i < n; ++i) {
        s = (s * 0x5DEECE66Dull + 0xBull) & mask;
        const int32_t hi = (int32_t)(s >> 16);
        s = (s * 0x5DEECE66Dull + 0xBull) & mask;
        const int32_t lo = (int32_t)(s >> 16);
        out[i] = (int64_t)(((uint64_t)(int64_t)hi << 32) + (uint64_t)(int64_t)lo);
    }
}

/* MTUtils.splitMethod(m,k,n,cores) (utils/MTUtils.scala:150-175) with dimToSplit (:204-213) */
void mo_split_method(int64_t m, int64_t k, int64_t n, int cores, int out[3]) {
    int ms = 1, ks = 1, ns = 1;
    while (cores > 1 && m > 1 && k > 1 && n > 1) {
        int d;
        if (n >= k && n >= m) d = 1; else if (m >= k && m >= n) d = 2; else d = 3;
        if (d == 1) { ns *= 2; n /= 2; } else if (d == 2) { ms *= 2; m /= 2; } else { ks *= 2; k /= 2; }
        cores /= 2;
    }
    out[0] = ms; out[1] = ks; out[2] = ns;
}
