/* The boundary from plain C: no Python, no torch -- include/td_b200.h and libtd_b200.so only.
 *
 *   gcc -I include -o /tmp/c_abi_demo examples/c_abi_demo.c -L gym_td_b200 -ltd_b200 \
 *       -Wl,-rpath,$PWD/gym_td_b200 -L/usr/local/cuda/lib64 -lcudart
 *   /tmp/c_abi_demo [n_envs] [steps]
 *
 * Generates maps on the host, creates a batch of TD-def-small envs, steps them with NOP / random actions and
 * prints the episode statistics.  Without a CUDA device td_create fails with TD_E_CUDA and a message (there is
 * no CPU fallback); the host-only entry points (td_mapgen*) still work.
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include "td_b200.h"

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 4096, steps = argc > 2 ? atoi(argv[2]) : 200, L = 10;
    printf("ABI version %d\n", td_abi_version());

    td_map *maps = (td_map *)calloc(n, sizeof(td_map));
    uint32_t *seeds = (uint32_t *)malloc(n * sizeof(uint32_t));
    int32_t *valid = (int32_t *)malloc(n * sizeof(int32_t));
    for (int i = 0; i < n; ++i) seeds[i] = 1000u + i;
    if (td_mapgen_batch(seeds, n, L, 0, 0, 1, 0, maps, valid) < 0) { fprintf(stderr, "mapgen failed\n"); return 1; }
    printf("map 0: seed %u, %d road(s), end cell %d, %d randint calls\n", seeds[0], maps[0].num_roads, maps[0].end,
           maps[0].n_randint);

    td_handle *h = NULL;
    int rc = td_create(NULL, TD_KIND_DEF, L, n, 0, &h);
    if (rc != TD_OK) {
        printf("td_create: %d (%s)\n", rc, td_last_error(NULL));
        return rc == TD_E_CUDA ? 3 : 1;                     /* 3: no GPU here -- expected on a CPU box */
    }
    float *obs; int64_t *act; double *reward; uint8_t *done;
    cudaMalloc((void **)&obs, (size_t)n * TD_NCHANNELS * L * L * sizeof(float));
    cudaMalloc((void **)&act, (size_t)n * sizeof(int64_t));
    cudaMalloc((void **)&reward, (size_t)n * sizeof(double));
    cudaMalloc((void **)&done, (size_t)n);
    int64_t *act_host = (int64_t *)malloc(n * sizeof(int64_t));
    uint32_t *opp = (uint32_t *)malloc(n * sizeof(uint32_t));
    for (int i = 0; i < n; ++i) opp[i] = seeds[i];
    if (td_upload_maps(h, maps, n) || td_seed_opponent_python(h, opp, 0, n) || td_reset(h, NULL, NULL, obs, NULL)) {
        fprintf(stderr, "setup: %s\n", td_last_error(h));
        return 1;
    }
    td_step_io io = {0};
    io.def_action_dev = act; io.auto_reset = 1; io.obs_dev = obs; io.reward_dev = reward; io.done_dev = done;
    io.obs_incremental = 1;                                  /* obs is only ever written by this handle */
    srand(1);
    for (int s = 0; s < steps; ++s) {
        for (int i = 0; i < n; ++i) act_host[i] = rand() % (6 * L * L + 1);
        cudaMemcpy(act, act_host, n * sizeof(int64_t), cudaMemcpyHostToDevice);
        if (td_step(h, &io, NULL)) { fprintf(stderr, "step: %s\n", td_last_error(h)); return 1; }
    }
    td_stats st;
    if (td_get_stats(h, &st, NULL)) { fprintf(stderr, "stats: %s\n", td_last_error(h)); return 1; }
    printf("%lld env-steps, %lld episodes finished, mean return %.3f\n", (long long)st.steps, (long long)st.episodes,
           st.episodes ? st.return_sum / (double)st.episodes : 0.0);
    td_destroy(h);
    return 0;
}
