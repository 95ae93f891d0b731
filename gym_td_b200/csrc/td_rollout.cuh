// td_rollout.cuh -- the rollout consumer around the env step, on device.
//
// What the reference's training loop does on the host, per vector-env step and per horizon:
//   * AllowNextMove masking: envs that may not move get empty_action()      train/main.py:130-132
//   * RealAction penalty: reward -= 0.3 when the executed action differs    train/PPO/Callbacks.py:21-23
//   * record(actions, rewards, dones) into [horizon, num_actors] buffers     train/PPO/Model.py:134-140
//   * flush: GAE(gamma, lam) advantages and returns                         train/PPO/Model.py:166-192
// Here the buffers live in HBM next to the envs ([horizon, n] layout, env index fastest), one thread per env.
// The GAE arithmetic follows the reference's NumPy evaluation (NEP 50 promotion): float32 buffers,
// float64 accumulation, float32 stores -- see oracle/rollout_oracle.py for the statement it is tested against.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace td {

// actions[i] <- empty action where the env may not move.  width = elements per env (1 Discrete, 24 cluster).
__global__ void rollout_mask_kernel(int64_t *actions, const uint8_t *allow, int n, int width, int allow_bit,
                                    int64_t empty_value)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * width) return;
    if (!(allow[i / width] & allow_bit)) actions[i] = empty_value;
}

// row t of the rollout buffers: penalised reward (float32), done, action.
__global__ void rollout_record_kernel(const int64_t *actions, const int64_t *real, const double *reward,
                                      const uint8_t *done, int n, int width, double penalty,
                                      float *rewards_row, uint8_t *dones_row, int64_t *actions_row)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool differs = false;
    for (int k = 0; k < width; ++k) {
        const int64_t a = actions[(size_t)i * width + k];
        differs = differs || (a != real[(size_t)i * width + k]);
        if (actions_row) actions_row[(size_t)i * width + k] = a;
    }
    double r = reward[i];
    if (differs) r = __dsub_rn(r, penalty);
    rewards_row[i] = (float)r;
    dones_row[i] = done[i];
}

// GAE over one horizon, reverse scan per env (train/PPO/Model.py:177-190).
__global__ void gae_kernel(int horizon, int n, const float *rewards, const uint8_t *dones, const float *values,
                           const float *next_value, double gamma, double lam, float *advs, float *returns)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gamma32 = (float)gamma;
    const double gl = __dmul_rn(gamma, lam);
    double last_gae = 0.0;
    for (int j = horizon - 1; j >= 0; --j) {
        const size_t at = (size_t)j * n + i;
        const double nn = 1.0 - (dones[at] ? 1.0 : 0.0);
        double g;
        if (j == horizon - 1) g = __dmul_rn(gamma, (double)next_value[i]);              // Python float * float
        else g = (double)__fmul_rn(gamma32, values[at + n]);                             // float * float32 array
        const double v = (double)values[at];
        const double delta = __dsub_rn(__dadd_rn((double)rewards[at], __dmul_rn(g, nn)), v);
        last_gae = __dadd_rn(delta, __dmul_rn(__dmul_rn(gl, nn), last_gae));
        const float a32 = (float)last_gae;
        advs[at] = a32;
        returns[at] = __fadd_rn(a32, values[at]);
    }
}

} // namespace td
