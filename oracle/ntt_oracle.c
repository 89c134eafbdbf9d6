/*
 * oracle/ntt_oracle.c -- CPU restatement of the GPU-NTT reference algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing on the product path (gpu_ntt_b200/csrc,
 * include/) may include, link or call this file.  Only tests/, the smoke check
 * in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py load the shared object built from it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 * (a) against oracle/_ref (the reference's own CPU sources compiled in place
 * from /root/reference by oracle/Makefile) whenever that library exists, and
 * (b) against the committed fixtures in tests/golden/ that were generated from
 * oracle/_ref by tests/golden/make_golden.py.
 *
 * Plain C11, gcc, unsigned __int128 for the 128-bit products.  All residues are
 * canonical (in [0,p)), so equality with the reference is exact equality.
 * Both element widths of the reference (Data32 / Data64) are carried in
 * uint64_t here; `width` (32 or 64) only selects the default parameter pools.
 *
 * Each function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef unsigned __int128 u128;

/* ---- Modulus<T>{value,bit,mu}: src/include/gpuntt/common/modular_arith.cuh:28-57
 * bit = (T)(log2(value) + 1) evaluated in DOUBLE precision exactly as the reference does
 * (:46) -- so a value within 2^-53 relative of a power of two (e.g. 2^61-1) gets the same
 * off-by-one bit count the reference computes; mu = floor(2^(2*bit+1) / value) truncated
 * to T (:49-56). */
void ora_modulus(uint64_t value, int width, uint64_t *bit, uint64_t *mu)
{
    uint64_t b = (uint64_t)(log2((double)value) + 1);
    *bit = b;
    if (width == 32) {
        uint64_t m = ((uint64_t)1 << (2 * b + 1)) / value;
        *mu = (uint32_t)m;
    } else {
        u128 m = ((u128)1 << (2 * b + 1)) / value;
        *mu = (uint64_t)m;
    }
}

/* ---- host Barrett multiply, restated literally:
 * src/include/gpuntt/common/modular_arith.cuh:91-108 (64-bit T2 for Data32,
 * 128-bit T2 for Data64).  Used by the tests to show it agrees with the plain
 * `%` product below for the supported modulus sizes. */
uint64_t ora_barrett_mult(uint64_t a, uint64_t b, uint64_t value, uint64_t bit,
                          uint64_t mu, int width)
{
    if (width == 32) {
        uint64_t mult = a * b;
        uint64_t r = mult >> (bit - 2);
        r = r * mu;
        r = r >> (bit + 3);
        r = r * value;
        mult = mult - r;
        uint32_t res = (uint32_t)(mult & 0xffffffffu);
        return (res >= value) ? (res - (uint32_t)value) : res;
    } else {
        u128 mult = (u128)a * b;
        u128 r = mult >> (bit - 2);
        r = r * (u128)mu;
        r = r >> (bit + 3);
        r = r * (u128)value;
        mult = mult - r;
        uint64_t res = (uint64_t)mult;
        return (res >= value) ? (res - value) : res;
    }
}

/* canonical a*b mod p -- what OPERATOR<T>::mult returns for reduced inputs */
static inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t p)
{
    return (uint64_t)(((u128)a * b) % p);
}
/* modular_arith.cuh:70-75 */
static inline uint64_t addmod(uint64_t a, uint64_t b, uint64_t p)
{
    uint64_t s = a + b;
    return (s >= p) ? s - p : s;
}
/* modular_arith.cuh:79-85 */
static inline uint64_t submod(uint64_t a, uint64_t b, uint64_t p)
{
    uint64_t d = a + p - b;
    return (d >= p) ? d - p : d;
}

uint64_t ora_mulmod(uint64_t a, uint64_t b, uint64_t p) { return mulmod(a, b, p); }

/* modular_arith.cuh:112-130 (square-and-multiply from the top bit) */
uint64_t ora_expmod(uint64_t base, uint64_t e, uint64_t p)
{
    uint64_t r = 1;
    if (e == 0) return r;
    int nb = 0;
    for (uint64_t t = e; t; t >>= 1) nb++;
    for (int i = nb - 1; i >= 0; i--) {
        r = mulmod(r, r, p);
        if ((e >> i) & 1) r = mulmod(r, base, p);
    }
    return r;
}
/* modular_arith.cuh:134-138: Fermat inverse */
uint64_t ora_modinv(uint64_t a, uint64_t p) { return ora_expmod(a, p - 2, p); }

/* src/lib/common/nttparameters.cu:10-20 */
int ora_bitreverse(int index, int n_power)
{
    int r = 0;
    for (int i = 0; i < n_power; i++) {
        r <<= 1;
        r = (index & 1) | r;
        index >>= 1;
    }
    return r;
}

/* ---- NTTParameters<T>(LOGN, poly) default pools:
 * src/lib/common/nttparameters.cu:22-49 (ctor), :84-98 (modulus_pool),
 * :100-120 (omega_pool), :122-142 (psi_pool), :170-173 (n_inverse).
 * poly: 0 = X_N_plus, 1 = X_N_minus (nttparameters.cuh:32-36).
 * out[0..7] = modulus, omega, psi, n_inv, root_of_unity,
 *             inverse_root_of_unity, root_of_unity_size, n */
void ora_merge_params(int logn, int poly, int width, uint64_t out[8])
{
    uint64_t p, omega, psi;
    if (width == 32) {
        p = 469762049ull;
        omega = ora_expmod(900, (uint32_t)(1u << (25 - logn)), p);
        psi = ora_expmod(30, (uint32_t)(1u << (25 - logn)), p);
    } else {
        p = 576460756061519873ull;
        /* the exponent is the int expression 1 << (28 - logn) (:116,:138) */
        omega = ora_expmod(229929041166717729ull, (uint64_t)(1 << (28 - logn)), p);
        psi = ora_expmod(4517306222ull, (uint64_t)(1 << (28 - logn)), p);
    }
    uint64_t n = (uint64_t)1 << logn;
    uint64_t root = (poly == 1) ? omega : psi;
    out[0] = p;
    out[1] = omega;
    out[2] = psi;
    out[3] = ora_modinv(n, p);
    out[4] = root;
    out[5] = ora_modinv(root, p);
    out[6] = (poly == 1) ? (n >> 1) : n;
    out[7] = n;
}

/* nttparameters.cu:144-168: table[i] = root^i, natural order */
void ora_power_table(uint64_t root, uint64_t p, uint64_t size, uint64_t *out)
{
    if (size == 0) return;
    out[0] = 1;
    for (uint64_t i = 1; i < size; i++) out[i] = mulmod(out[i - 1], root, p);
}

/* nttparameters.cu:175-189 / :455-469: out[i] = table[bitreverse(i, log2 size)] */
void ora_bitrev_table(const uint64_t *table, uint64_t size, uint64_t *out)
{
    int lg = 0;
    while (((uint64_t)1 << lg) < size) lg++;
    for (uint64_t i = 0; i < size; i++) out[i] = table[ora_bitreverse((int)i, lg)];
}

/* ---- NTTCPU<T>::ntt: src/lib/ntt_merge/ntt_cpu.cu:81-128.
 * In place on a[0..n), natural-order input, bit-reversed-order output.
 * fwd_table is the NATURAL-order power table (forward_root_of_unity_table). */
void ora_merge_ntt(uint64_t *a, int logn, uint64_t p, const uint64_t *fwd_table, int poly)
{
    int n = 1 << logn;
    int t = n, m = 1;
    while (m < n) {
        t >>= 1;
        for (int i = 0; i < m; i++) {
            int j1 = 2 * i * t, j2 = j1 + t - 1;
            int index = (poly == 1) ? ora_bitreverse(i, logn - 1) : ora_bitreverse(m + i, logn);
            uint64_t S = fwd_table[index];
            for (int j = j1; j <= j2; j++) {
                uint64_t U = a[j];
                uint64_t V = mulmod(a[j + t], S, p);
                a[j] = addmod(U, V, p);
                a[j + t] = submod(U, V, p);
            }
        }
        m <<= 1;
    }
}

/* ---- NTTCPU<T>::intt: src/lib/ntt_merge/ntt_cpu.cu:130-185 (GS stages then
 * a multiply by n^-1).  inv_table natural order. */
void ora_merge_intt(uint64_t *a, int logn, uint64_t p, const uint64_t *inv_table, int poly)
{
    int n = 1 << logn;
    int t = 1, m = n;
    while (m > 1) {
        int j1 = 0, h = m >> 1;
        for (int i = 0; i < h; i++) {
            int j2 = j1 + t - 1;
            int index = (poly == 1) ? ora_bitreverse(i, logn - 1) : ora_bitreverse(h + i, logn);
            uint64_t S = inv_table[index];
            for (int j = j1; j <= j2; j++) {
                uint64_t U = a[j], V = a[j + t];
                a[j] = addmod(U, V, p);
                a[j + t] = mulmod(submod(U, V, p), S, p);
            }
            j1 += (t << 1);
        }
        t <<= 1;
        m >>= 1;
    }
    uint64_t n_inv = ora_modinv((uint64_t)n, p);
    for (int i = 0; i < n; i++) a[i] = mulmod(a[i], n_inv, p);
}

/* signed-input reduction on the forward path:
 * src/include/gpuntt/common/modular_arith.cuh:372-385 (used ntt.cu:481-489) */
uint64_t ora_reduce_signed(int64_t x, uint64_t p)
{
    return (x < 0) ? p - (uint64_t)(-x) : (uint64_t)x;
}
/* signed centred output on the inverse path: modular_arith.cuh:389-405
 * (used ntt.cu:1178-1186) */
int64_t ora_centered(uint64_t x, uint64_t p)
{
    uint64_t half = p >> 1;
    return (x > half) ? (int64_t)x - (int64_t)p : (int64_t)x;
}

/* ---- schoolbook_poly_multiplication: ntt_cpu.cu:10-57 (O(n^2)) */
void ora_schoolbook(const uint64_t *a, const uint64_t *b, int n, uint64_t p, int poly,
                    uint64_t *out)
{
    uint64_t *mv = (uint64_t *)calloc((size_t)2 * n, sizeof(uint64_t));
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            mv[i + j] = addmod(mv[i + j], mulmod(a[i], b[j], p), p);
    for (int i = 0; i < n; i++)
        out[i] = (poly == 1) ? addmod(mv[i], mv[i + n], p) : submod(mv[i], mv[i + n], p);
    free(mv);
}

/* =======================  4-step  ======================= */

/* NTTParameters4Step<T> pools: nttparameters.cu:229-303 (indexed logn-12) */
static const uint64_t P4_64[] = {
    576460752303415297ull, 576460752303439873ull, 576460752304439297ull,
    576460752308273153ull, 576460752308273153ull, 576460752315482113ull,
    576460752315482113ull, 576460752340123649ull, 576460752364240897ull,
    576460752475389953ull, 576460752597024769ull, 576460753024843777ull,
    576460753175838721ull};
static const uint64_t W4_64[] = {
    288482366111684746ull, 37048445140799662ull,  459782973201979845ull,
    64800917766465203ull,  425015386842055933ull, 18734847765732801ull,
    119109113519742895ull, 227584740857897520ull, 477282059544659462ull,
    570131728462077067ull, 433594414095420776ull, 219263994987749328ull,
    189790554094222112ull};
static const uint64_t PSI4_64[] = {
    238394956950829ull, 54612008597396ull, 8242615629351ull, 16141297350887ull,
    3760097055997ull,   11571974431275ull, 328867687796ull,  2298846063117ull,
    731868219707ull,    409596963254ull,   189266227206ull,  31864818375ull,
    92067739764ull};
static const uint64_t P4_32[] = {268460033, 268582913, 268664833, 268369921, 269221889,
                                 269221889, 270532609, 270532609, 270532609, 377487361,
                                 377487361, 469762049, 469762049};
static const uint64_t W4_32[] = {36747374, 249229369, 4092529, 175218169, 10653696, 238764304,
                                 240100,   23104,     179776,  19321,     38809,    1600,
                                 169};
static const uint64_t PSI4_32[] = {77090, 15787, 2023, 13237, 3264, 15452, 490,
                                   152,   424,   139,  197,   40,   13};

/* matrix_dimention(): nttparameters.cu:305-354 */
static const int N1_4[] = {32, 32, 32, 64, 128, 32, 32, 32, 32, 64, 128, 128, 256};
static const int N2_4[] = {128, 256, 512, 512, 512, 4096, 8192, 16384, 32768, 32768, 32768, 65536,
                           65536};

/* NTTParameters4Step ctor: nttparameters.cu:191-225.
 * out[0..9] = modulus, omega, psi, n_inv, root, inv_root, root_size, n, n1, n2 */
int ora_4step_params(int logn, int poly, int width, uint64_t out[10])
{
    if (logn < 12 || logn > 24) return -1;
    int k = logn - 12;
    uint64_t p = (width == 32) ? P4_32[k] : P4_64[k];
    uint64_t omega = (width == 32) ? W4_32[k] : W4_64[k];
    uint64_t psi = (width == 32) ? PSI4_32[k] : PSI4_64[k];
    uint64_t n = (uint64_t)1 << logn;
    uint64_t root = (poly == 1) ? omega : psi;
    out[0] = p;
    out[1] = omega;
    out[2] = psi;
    out[3] = ora_modinv(n, p);
    out[4] = root;
    out[5] = ora_modinv(root, p);
    out[6] = (poly == 1) ? (n >> 1) : n;
    out[7] = n;
    out[8] = (uint64_t)N1_4[k];
    out[9] = (uint64_t)N2_4[k];
    return 0;
}

/* small_{forward,inverse}_root_of_unity_table_generator: nttparameters.cu:356-380,
 * :398-427.  Natural-order tables of n1/2 and n2/2 powers of root^(n/n1), root^(n/n2)
 * (inverse: of their modular inverses). */
void ora_4step_small_tables(uint64_t root, uint64_t p, uint64_t n, int n1, int n2, int inverse,
                            uint64_t *t1, uint64_t *t2)
{
    uint64_t r1 = ora_expmod(root, n / (uint64_t)n1, p);
    uint64_t r2 = ora_expmod(root, n / (uint64_t)n2, p);
    if (inverse) {
        r1 = ora_modinv(r1, p);
        r2 = ora_modinv(r2, p);
    }
    ora_power_table(r1, p, (uint64_t)(n1 >> 1), t1);
    ora_power_table(r2, p, (uint64_t)(n2 >> 1), t2);
}

/* TW_forward_table_generator: nttparameters.cu:382-396
 *   W[i*n2+j] = root^(bitreverse(i, log2 n1) * j)
 * TW_inverse_table_generator: nttparameters.cu:429-443
 *   Winv[i*n2+j] = inv_root^(bitreverse(j, log2 n2) * i)
 * (computed here with running products instead of one exp per entry; the
 * residues are identical). */
void ora_4step_w_table(uint64_t root_or_invroot, uint64_t p, int n1, int n2, int inverse,
                       uint64_t *W)
{
    int lg1 = 0, lg2 = 0;
    while ((1 << lg1) < n1) lg1++;
    while ((1 << lg2) < n2) lg2++;
    if (!inverse) {
        for (int i = 0; i < n1; i++) {
            uint64_t g = ora_expmod(root_or_invroot, (uint64_t)ora_bitreverse(i, lg1), p);
            uint64_t acc = 1;
            for (int j = 0; j < n2; j++) {
                W[(size_t)i * n2 + j] = acc;
                acc = mulmod(acc, g, p);
            }
        }
    } else {
        /* column j holds powers (in i) of inv_root^bitreverse(j) */
        uint64_t *g = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n2);
        for (int j = 0; j < n2; j++)
            g[j] = ora_expmod(root_or_invroot, (uint64_t)ora_bitreverse(j, lg2), p);
        for (int j = 0; j < n2; j++) W[j] = 1;
        for (int i = 1; i < n1; i++)
            for (int j = 0; j < n2; j++)
                W[(size_t)i * n2 + j] = mulmod(W[(size_t)(i - 1) * n2 + j], g[j], p);
        free(g);
    }
}

/* core_ntt: src/lib/ntt_4step/ntt_4step_cpu.cu:111-147 (always the X^N-1 index rule) */
static void core_ntt(uint64_t *a, const uint64_t *tab, int lg, uint64_t p)
{
    int n = 1 << lg, t = n, m = 1;
    while (m < n) {
        t >>= 1;
        for (int i = 0; i < m; i++) {
            int j1 = 2 * i * t;
            uint64_t S = tab[ora_bitreverse(i, lg - 1)];
            for (int j = j1; j < j1 + t; j++) {
                uint64_t U = a[j], V = mulmod(a[j + t], S, p);
                a[j] = addmod(U, V, p);
                a[j + t] = submod(U, V, p);
            }
        }
        m <<= 1;
    }
}
/* core_intt: ntt_4step_cpu.cu:148-191 */
static void core_intt(uint64_t *a, const uint64_t *tab, int lg, uint64_t p)
{
    int n = 1 << lg, t = 1, m = n;
    while (m > 1) {
        int j1 = 0, h = m >> 1;
        for (int i = 0; i < h; i++) {
            uint64_t S = tab[ora_bitreverse(i, lg - 1)];
            for (int j = j1; j < j1 + t; j++) {
                uint64_t U = a[j], V = a[j + t];
                a[j] = addmod(U, V, p);
                a[j + t] = mulmod(submod(U, V, p), S, p);
            }
            j1 += (t << 1);
        }
        t <<= 1;
        m >>= 1;
    }
}

/* NTT_4STEP_CPU<T>::ntt: ntt_4step_cpu.cu:33-68.
 * in/out: n = n1*n2 elements; t1/t2 natural-order small tables; W = forward W table. */
void ora_4step_ntt(const uint64_t *in, uint64_t *out, int n1, int n2, uint64_t p,
                   const uint64_t *t1, const uint64_t *t2, const uint64_t *W)
{
    int lg1 = 0, lg2 = 0;
    while ((1 << lg1) < n1) lg1++;
    while ((1 << lg2) < n2) lg2++;
    size_t n = (size_t)n1 * n2;
    uint64_t *T = (uint64_t *)malloc(sizeof(uint64_t) * n);
    uint64_t *B = (uint64_t *)malloc(sizeof(uint64_t) * n);
    /* matrix n1 x n2 -> transpose n2 x n1 (:36-39) */
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n2; j++) T[(size_t)j * n1 + i] = in[(size_t)i * n2 + j];
    for (int j = 0; j < n2; j++) core_ntt(T + (size_t)j * n1, t1, lg1, p); /* :41-46 */
    /* transpose back to n1 x n2 and multiply by W elementwise (:48-52) */
    for (int j = 0; j < n2; j++)
        for (int i = 0; i < n1; i++)
            B[(size_t)i * n2 + j] = mulmod(T[(size_t)j * n1 + i], W[(size_t)i * n2 + j], p);
    for (int i = 0; i < n1; i++) core_ntt(B + (size_t)i * n2, t2, lg2, p); /* :54-62 */
    /* final transpose to n2 x n1 (:64-65) */
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n2; j++) out[(size_t)j * n1 + i] = B[(size_t)i * n2 + j];
    free(T);
    free(B);
}

/* NTT_4STEP_CPU<T>::intt: ntt_4step_cpu.cu:70-109; vector_to_matrix_intt :227-243
 * (n2 rows of n1 elements: the flat input is consumed n1 elements at a time in the
 * order array[i + j*rows], i.e. it is read as the row-major n2 x n1 matrix already). */
void ora_4step_intt(const uint64_t *in, uint64_t *out, int n1, int n2, uint64_t p,
                    const uint64_t *t1inv, const uint64_t *t2inv, const uint64_t *Winv,
                    uint64_t n_inv)
{
    int lg1 = 0, lg2 = 0;
    while ((1 << lg1) < n1) lg1++;
    while ((1 << lg2) < n2) lg2++;
    size_t n = (size_t)n1 * n2;
    uint64_t *T = (uint64_t *)malloc(sizeof(uint64_t) * n);
    uint64_t *B = (uint64_t *)malloc(sizeof(uint64_t) * n);
    /* vector_to_matrix_intt(input, rows=n1, cols=n2): element (i,j) = array[i + j*n1]
     * is appended to row floor((i*n2+j)/n1) (:233-240) */
    {
        size_t *fill = (size_t *)calloc((size_t)n2, sizeof(size_t));
        for (int i = 0; i < n1; i++)
            for (int j = 0; j < n2; j++) {
                size_t r = ((size_t)i * n2 + j) / (size_t)n1;
                T[r * n1 + fill[r]++] = in[(size_t)i + (size_t)j * n1];
            }
        free(fill);
    }
    for (int j = 0; j < n2; j++) core_intt(T + (size_t)j * n1, t1inv, lg1, p);
    for (int j = 0; j < n2; j++)
        for (int i = 0; i < n1; i++)
            B[(size_t)i * n2 + j] = mulmod(T[(size_t)j * n1 + i], Winv[(size_t)i * n2 + j], p);
    for (int i = 0; i < n1; i++) core_intt(B + (size_t)i * n2, t2inv, lg2, p);
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n2; j++)
            out[(size_t)j * n1 + i] = mulmod(B[(size_t)i * n2 + j], n_inv, p);
    free(T);
    free(B);
}

/* intt_first_transpose: ntt_4step_cpu.cu:287-299 = flatten(vector_to_matrix_intt(in)) */
void ora_4step_intt_first_transpose(const uint64_t *in, uint64_t *out, int n1, int n2)
{
    size_t *fill = (size_t *)calloc((size_t)n2, sizeof(size_t));
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n2; j++) {
            size_t r = ((size_t)i * n2 + j) / (size_t)n1;
            out[r * n1 + fill[r]++] = in[(size_t)i + (size_t)j * n1];
        }
    free(fill);
}

/* fold hash used by the fixtures: h = h*1000003 + v (mod 2^64) */
uint64_t ora_fold_hash(const uint64_t *v, size_t n)
{
    uint64_t h = 0;
    for (size_t i = 0; i < n; i++) h = h * 1000003ull + v[i];
    return h;
}
