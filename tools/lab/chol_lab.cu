// chol_lab.cu -- lab harness for blu_kkt_chol_kernel (blocked one-CTA Cholesky + solve of the KKT capacitance matrix).
// Builds a random SPD matrix of order Q with a right-hand side, runs the kernel on the column-major + rhs-row layout the
// capfold kernel produces, compares y with a host Cholesky solve, prints the event time.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Iinclude -Ibluest_b200/csrc -o tools/lab/bin/chol_lab tools/lab/chol_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#define BLU_CHOL_STAMPS
#include "blu_kkt.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static void run(int Q)
{
    const int LD = ((Q + 1 + 3) / 4) * 4, K = Q + 40;
    std::vector<double> B((size_t)K * Q), A((size_t)Q * Q), v(Q), cap((size_t)Q * LD, 0.0);
    for (auto &x : B) x = (rand() / (double)RAND_MAX) - 0.5;
    for (auto &x : v) x = (rand() / (double)RAND_MAX) - 0.5;
    for (int i = 0; i < Q; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = 0; k < K; ++k) s += B[(size_t)k * Q + i] * B[(size_t)k * Q + j];
            A[(size_t)i * Q + j] = A[(size_t)j * Q + i] = s;
        }
    for (int i = 0; i < Q; ++i) for (int j = 0; j < Q; ++j) cap[(size_t)j * LD + i] = A[(size_t)i * Q + j];
    for (int j = 0; j < Q; ++j) cap[(size_t)j * LD + Q] = v[j];
    // host solve
    std::vector<double> Lh(A), yh(v);
    for (int j = 0; j < Q; ++j) {
        double d = Lh[(size_t)j * Q + j];
        for (int k = 0; k < j; ++k) d -= Lh[(size_t)j * Q + k] * Lh[(size_t)j * Q + k];
        d = sqrt(d);
        Lh[(size_t)j * Q + j] = d;
        for (int i = j + 1; i < Q; ++i) {
            double s = Lh[(size_t)i * Q + j];
            for (int k = 0; k < j; ++k) s -= Lh[(size_t)i * Q + k] * Lh[(size_t)j * Q + k];
            Lh[(size_t)i * Q + j] = s / d;
        }
    }
    for (int i = 0; i < Q; ++i) { double s = yh[i]; for (int k = 0; k < i; ++k) s -= Lh[(size_t)i * Q + k] * yh[k]; yh[i] = s / Lh[(size_t)i * Q + i]; }
    for (int i = Q - 1; i >= 0; --i) { double s = yh[i]; for (int k = i + 1; k < Q; ++k) s -= Lh[(size_t)k * Q + i] * yh[k]; yh[i] = s / Lh[(size_t)i * Q + i]; }
    double *d_cap0, *d_cap, *d_y; int *d_info;
    CK(cudaMalloc(&d_cap0, sizeof(double) * cap.size())); CK(cudaMalloc(&d_cap, sizeof(double) * cap.size()));
    CK(cudaMalloc(&d_y, sizeof(double) * 256)); CK(cudaMalloc(&d_info, 4));
    CK(cudaMemcpy(d_cap0, cap.data(), sizeof(double) * cap.size(), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0.f;
    const int reps = 10;
    for (int it = 0; it < reps + 2; ++it) {
        CK(cudaMemcpy(d_cap, d_cap0, sizeof(double) * cap.size(), cudaMemcpyDeviceToDevice));
        CK(cudaMemset(d_info, 0xff, 4));
        CK(cudaEventRecord(e0));
        blu_kkt_chol_kernel<<<1, BLU_CHOL_T>>>(d_cap, Q, LD, d_y, d_info);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 2) { sum += ms; if (ms < best) best = ms; }
    }
    long long cyc[8];
    CK(cudaMemcpyFromSymbol(cyc, blu_chol_cycles, sizeof(cyc)));
    printf("         cycles per run: panel load %lld, diagonal block %lld, rows below %lld, panel store %lld, trailing update %lld, backward solve %lld, backward update %lld\n",
           cyc[0] / (reps + 2), cyc[1] / (reps + 2), cyc[2] / (reps + 2), cyc[3] / (reps + 2), cyc[4] / (reps + 2), cyc[5] / (reps + 2), cyc[6] / (reps + 2));
    { long long z[8] = {0}; CK(cudaMemcpyToSymbol(blu_chol_cycles, z, sizeof(z))); }
    std::vector<double> y(Q); int info = -2;
    CK(cudaMemcpy(y.data(), d_y, sizeof(double) * Q, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&info, d_info, 4, cudaMemcpyDeviceToHost));
    double err = 0.0, mx = 0.0;
    for (int i = 0; i < Q; ++i) { err = fmax(err, fabs(y[i] - yh[i])); mx = fmax(mx, fabs(yh[i])); }
    printf("Q = %3d: info %d, %.1f us mean, %.1f us best, max |y - y_host| / max |y_host| = %.2e\n", Q, info, sum / reps * 1e3, best * 1e3, err / mx);
    // a matrix that is not positive definite must be reported, not hang
    cap[(size_t)(Q / 2) * LD + Q / 2] = -1.0;
    CK(cudaMemcpy(d_cap, cap.data(), sizeof(double) * cap.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_info, 0xff, 4));
    blu_kkt_chol_kernel<<<1, BLU_CHOL_T>>>(d_cap, Q, LD, d_y, d_info);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&info, d_info, 4, cudaMemcpyDeviceToHost));
    printf("         indefinite matrix: info %d (expected %d)\n", info, Q / 2 + 1);
    cudaFree(d_cap0); cudaFree(d_cap); cudaFree(d_y); cudaFree(d_info);
}

int main()
{
    srand(3);
    const int qs[] = {5, 16, 17, 33, 138, 173, 233, 255};
    for (int q : qs) run(q);
    return 0;
}
