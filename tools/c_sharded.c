/* tools/c_sharded.c -- a plain-C host running the sharded entry points (BASELINE config 5 in miniature): no Python, no torch,
 * no nccl.h.  One process per GPU, launched by anything that sets RANK / WORLD_SIZE / LOCAL_RANK (e.g.
 *   python -m torch.distributed.run --no-python --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/c_sharded ).
 * Rank 0 obtains the NCCL unique id through the library and publishes it in a file; every rank shards a deterministic
 * job, calls c25519_x25519_shared_sharded / c25519_ed25519_verify_sharded and checks the GATHERED arrays against the
 * whole job computed locally with the plain batch entry points (every rank can do that: the inputs are deterministic).
 * The X25519 call runs twice: on an unregistered array (NCCL all-gather) and on a registered one (fused peer-memory path).
 * Build: gcc -O2 -Iinclude -I/usr/local/cuda/include tools/c_sharded.c -Lcurve25519_b200 -lcurve25519_b200 -L/usr/local/cuda/lib64 -lcudart
 *        -Wl,-rpath,'$ORIGIN/../curve25519_b200' -o tools/c_sharded */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <cuda_runtime_api.h>
#include "c25519_b200.h"

#define CHECK(x) do { int rc__ = (x); if (rc__) { fprintf(stderr, "rank %d: %s failed (%d): %s\n", rank, #x, rc__, c25519_last_error()); return 1; } } while (0)
#define CUCHECK(x) do { cudaError_t e__ = (x); if (e__) { fprintf(stderr, "rank %d: %s: %s\n", rank, #x, cudaGetErrorString(e__)); return 1; } } while (0)

static int envi(const char *k, int d) { const char *v = getenv(k); return v ? atoi(v) : d; }
static void fill(unsigned char *p, size_t n, unsigned long long s)
{ for (size_t i = 0; i < n; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; p[i] = (unsigned char)(s >> 24); } }

int main(void)
{
    const int rank = envi("RANK", 0), world = envi("WORLD_SIZE", 1), local = envi("LOCAL_RANK", 0);
    const size_t n_local = 40000, n = n_local * (size_t)world;
    char path[128]; snprintf(path, sizeof path, "/tmp/c25519_uid_%s", getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0");
    unsigned char id[C25519_NCCL_UNIQUE_ID_BYTES];
    CUCHECK(cudaSetDevice(local));
    CHECK(c25519_init(local));
    if (rank == 0) {
        CHECK(c25519_nccl_unique_id(id));
        char tmp[160]; snprintf(tmp, sizeof tmp, "%s.tmp", path);
        FILE *f = fopen(tmp, "wb"); fwrite(id, 1, sizeof id, f); fclose(f); rename(tmp, path);
    } else {
        FILE *f = NULL;
        for (int t = 0; t < 600 && !(f = fopen(path, "rb")); t++) usleep(100000);
        if (!f || fread(id, 1, sizeof id, f) != sizeof id) { fprintf(stderr, "rank %d: no unique id\n", rank); return 1; }
        fclose(f);
    }
    void *comm = NULL;
    CHECK(c25519_nccl_comm_init(&comm, world, rank, id, local));
    /* the whole deterministic job on the host, this rank's shard + the full job on the device */
    unsigned char *h_sk = malloc(32 * n), *h_pk = malloc(32 * n), *h_all = malloc(32 * n), *h_exp = malloc(32 * n);
    fill(h_sk, 32 * n, 0x1234567ull); fill(h_pk, 32 * n, 0x7654321ull);
    unsigned char *d_sk, *d_pk, *d_sk_l, *d_pk_l, *d_all, *d_reg, *d_exp;
    CUCHECK(cudaMalloc((void **)&d_sk, 32 * n)); CUCHECK(cudaMalloc((void **)&d_pk, 32 * n)); CUCHECK(cudaMalloc((void **)&d_exp, 32 * n));
    CUCHECK(cudaMalloc((void **)&d_sk_l, 32 * n_local)); CUCHECK(cudaMalloc((void **)&d_pk_l, 32 * n_local));
    CUCHECK(cudaMalloc((void **)&d_all, 32 * n)); CUCHECK(cudaMalloc((void **)&d_reg, 32 * n));
    CUCHECK(cudaMemcpy(d_sk, h_sk, 32 * n, cudaMemcpyHostToDevice)); CUCHECK(cudaMemcpy(d_pk, h_pk, 32 * n, cudaMemcpyHostToDevice));
    CHECK(c25519_x25519_shared_batch(d_exp, d_pk, d_sk, n, NULL));                 /* expected: the whole job, locally */
    CUCHECK(cudaMemcpy(h_exp, d_exp, 32 * n, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int pass = 0; pass < 2; pass++) {
        unsigned char *dst = pass ? d_reg : d_all;
        if (pass) CHECK(c25519_sharded_register(d_reg, 32 * n, comm));             /* collective: fused peer-memory exchange from here on */
        CUCHECK(cudaMemcpy(d_sk_l, h_sk + 32 * n_local * rank, 32 * n_local, cudaMemcpyHostToDevice));
        CUCHECK(cudaMemcpy(d_pk_l, h_pk + 32 * n_local * rank, 32 * n_local, cudaMemcpyHostToDevice));
        CUCHECK(cudaMemset(dst, 0, 32 * n));
        CUCHECK(cudaDeviceSynchronize());
        CHECK(c25519_x25519_shared_sharded(dst, d_pk_l, d_sk_l, n_local, comm, NULL));
        CUCHECK(cudaDeviceSynchronize());
        CUCHECK(cudaMemcpy(h_all, dst, 32 * n, cudaMemcpyDeviceToHost));
        const int same = memcmp(h_all, h_exp, 32 * n) == 0;
        printf("rank %d/%d: %s gathered array %s the locally computed whole job (%zu shared keys)\n", rank, world,
               pass ? "registered (fused peer-memory)" : "NCCL all-gather:", same ? "==" : "!=", n);
        bad += !same;
    }
    CHECK(c25519_sharded_unregister(d_reg));
    CHECK(c25519_nccl_comm_destroy(comm));
    if (rank == 0) unlink(path);
    return bad;
}
