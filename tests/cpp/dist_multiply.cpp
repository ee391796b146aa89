// A compiled, torch-free rank program for the multi-GPU multiply: BlockMatrix.multiply(other, comm) of the C++ host mirror
// (include/marlin_b200.hpp) -> mb_matmul_blocked_dist.  Started once per rank (tests/test_gpu_dist_cabi.py):
//     dist_multiply <rank> <world> <session> <visible devices> [M K N m k n]
// Every rank builds the same A and B on the host (a fixed integer recurrence, values in [-4, 4] so every product and sum
// is exact in fp64), keeps the blocks MatrixElemOpPartitioner order mod world gives it, multiplies collectively, and
// compares the C blocks it ends up owning with the host product — exactly (small integers: no rounding anywhere).
// Exit status: 0 = ok, 1 = wrong result, 3 = no usable GPU / communicator.
#include "marlin_b200.hpp"

#include <cstdio>
#include <cstdlib>
#include <string>

using namespace marlin;

static DenseMatrix pattern(int rows, int cols, unsigned seed) {
    DenseMatrix d(rows, cols);
    unsigned s = seed;
    for (int c = 0; c < cols; ++c)
        for (int r = 0; r < rows; ++r) {
            s = s * 1664525u + 1013904223u;
            d(r, c) = (double)((int)((s >> 24) % 9u) - 4);
        }
    return d;
}
static DenseMatrix part(const DenseMatrix& a, int r0, int r1, int c0, int c1) {
    DenseMatrix d(r1 - r0, c1 - c0);
    for (int c = c0; c < c1; ++c)
        for (int r = r0; r < r1; ++r) d(r - r0, c - c0) = a(r, c);
    return d;
}

int main(int argc, char** argv) {
    if (argc < 5) { std::printf("usage: dist_multiply rank world session ndev [M K N m k n]\n"); return 2; }
    const int rank = std::atoi(argv[1]), world = std::atoi(argv[2]), ndev = std::atoi(argv[4]);
    const std::string session = argv[3];
    int dims[6] = {700, 530, 900, 2, 3, 2};
    for (int i = 0; i < 6 && 5 + i < argc; ++i) dims[i] = std::atoi(argv[5 + i]);
    const int M = dims[0], K = dims[1], N = dims[2], m = dims[3], k = dims[4], n = dims[5];
    try {
        Context::get(ndev > 0 ? rank % ndev : 0);
    } catch (const std::exception& e) {
        std::printf("cannot initialise marlin_b200: %s\n", e.what());
        return 3;
    }
    try {
        Comm comm(rank, world, session);
        const DenseMatrix A = pattern(M, K, 7u), B = pattern(K, N, 11u);
        const int rl = ceilLen(M, m), kl = ceilLen(K, k), cl = ceilLen(N, n);
        BlockMatrix::Blocks ba, bb;
        for (int i = 0; i < m; ++i)
            for (int kk = 0; kk < k; ++kk)
                if (comm.owner(i, kk, k) == rank)
                    ba.emplace_back(BlockID(i, kk), SubMatrix(part(A, i * rl, std::min(M, (i + 1) * rl), kk * kl, std::min(K, (kk + 1) * kl))));
        for (int kk = 0; kk < k; ++kk)
            for (int j = 0; j < n; ++j)
                if (comm.owner(kk, j, n) == rank)
                    bb.emplace_back(BlockID(kk, j), SubMatrix(part(B, kk * kl, std::min(K, (kk + 1) * kl), j * cl, std::min(N, (j + 1) * cl))));
        BlockMatrix a(ba, M, K, m, k), b(bb, K, N, k, n);
        int bad = 0, blocks = 0;
        for (int round = 0; round < 2; ++round) {             // twice: the second call reuses mappings, staging and flags
            BlockMatrix c = a.multiply(b, comm);
            for (auto& kv : c.blocks) {
                const DenseMatrix got = kv.second.denseBlock();
                const int r0 = kv.first.row * rl, c0 = kv.first.column * cl;
                for (int cc = 0; cc < got.cols; ++cc)
                    for (int rr = 0; rr < got.rows; ++rr) {
                        double want = 0.0;
                        for (int x = 0; x < K; ++x) want += A(r0 + rr, x) * B(x, c0 + cc);
                        if (got(rr, cc) != want) ++bad;
                    }
                ++blocks;
            }
            comm.checkPeers();
        }
        comm.barrier();
        if (bad) { std::printf("cpp rank %d/%d: %d wrong elements\n", rank, world, bad); return 1; }
        std::printf("cpp rank %d/%d ok (%d C blocks checked exactly)\n", rank, world, blocks);
        return 0;
    } catch (const std::exception& e) {
        std::printf("cpp rank %d/%d failed: %s\n", rank, world, e.what());
        return 3;
    }
}
