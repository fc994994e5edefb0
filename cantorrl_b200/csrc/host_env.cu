// Host-buffer face of the vectorised env: what a NumPy / SubprocVecEnv-style caller binds.
//
// The handle owns the device copies (book, state, action / obs / reward / done staging); the caller passes
// HOST pointers, ideally page-locked (cantor_host_register).  One step = for each of C chunks of envs:
// H2D of the chunk's actions -> fused hedge-step kernel on the chunk -> D2H of its obs / reward / done, the
// chunks round-robined over three streams so the copies of one chunk overlap the kernel and the opposite-
// direction copy of its neighbours (PCIe is full duplex).  The call returns when the results are on the host,
// because a gym caller needs this step's observation before it can choose the next action.
//
// Replaces: HedgingEnv.__init__ / reset / step (src/env/hedging_env_v2.py:10-77, :145-173, :175-294) as seen
// through SB3's VecEnv by src/agents/train_ppo_v2.py:127-141.
#include <chrono>
#include <new>
#include <vector>

#include "common.cuh"

extern "C" int cantor_env_step(const cantor_env_params*, const cantor_replay_book*, const cantor_env_state*, int64_t,
                               int32_t, const float*, float*, void*, uint8_t*, float*, int32_t,
                               const cantor_reset_rule*, const cantor_info_out*, void*);
extern "C" int cantor_env_reset(const cantor_env_params*, const cantor_replay_book*, const cantor_env_state*, int64_t,
                                int32_t, const uint8_t*, const int32_t*, float*, void*);
extern "C" int cantor_pack_book(const void*, const void*, const void*, const void*, int32_t, int32_t, int32_t, float*,
                                int64_t, void*);
extern "C" int cantor_sim_paths(const cantor_sim_params*, int32_t, int32_t, float*, int64_t, void*);

struct cantor_vecenv {
    cantor_env_params params;
    int precision, device, n_chunks;
    int64_t n_envs;
    // device memory
    float* book = nullptr;
    int64_t ld = 0;
    int32_t n_paths = 0, T = 0;
    int32_t* core = nullptr;
    void* cash = nullptr;
    double* pv_prev = nullptr;
    float* actions = nullptr;
    float* obs = nullptr;
    void* reward = nullptr;
    uint8_t* done = nullptr;
    int32_t* next_path = nullptr;
    cantor_reset_rule rule;
    int64_t global_step = 0;
    bool direct_small_outputs = true;     // CANTOR_HOST_NO_DIRECT=1 in the environment turns the mapped-memory path off (A/B runs)
    cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
};

namespace {

using namespace cantor;

size_t reward_bytes(const cantor_vecenv* e) { return e->precision == CANTOR_F64 ? sizeof(double) : sizeof(float); }

int free_all(cantor_vecenv* e) {
    cudaSetDevice(e->device);
    for (auto& s : e->streams) if (s) cudaStreamDestroy(s);
    cudaFree(e->book); cudaFree(e->core); cudaFree(e->cash); cudaFree(e->pv_prev); cudaFree(e->actions);
    cudaFree(e->obs); cudaFree(e->reward); cudaFree(e->done); cudaFree(e->next_path);
    delete e;
    return CANTOR_OK;
}

int alloc_book(cantor_vecenv* e, int32_t n_paths, int32_t T) {
    CANTOR_REQUIRE(n_paths > 0 && T > 0, "empty book");
    CANTOR_CUDA(cudaSetDevice(e->device));
    if (e->book) CANTOR_CUDA(cudaFree(e->book));
    e->book = nullptr;
    e->ld = (n_paths + 7) / 8 * 8;
    e->n_paths = n_paths;
    e->T = T;
    const size_t bytes = (size_t)(T + 1) * e->ld * 4 * sizeof(float);
    CANTOR_CUDA(cudaMalloc(&e->book, bytes));
    CANTOR_CUDA(cudaMemsetAsync(e->book, 0, bytes, e->streams[0]));
    return CANTOR_OK;
}

cantor_replay_book book_of(const cantor_vecenv* e) { return cantor_replay_book{e->book, e->ld, e->n_paths, e->T}; }

cantor_env_state state_of(const cantor_vecenv* e, int64_t first) {
    cantor_env_state st = {};                      // Monitor fields stay NULL on the host-buffer face
    st.core = e->core + first * 4;
    st.cash = (char*)e->cash + first * reward_bytes(e);
    st.pv_prev = e->pv_prev ? e->pv_prev + first : nullptr;
    return st;
}

}  // namespace

static bool env_flag(const char* name) { const char* v = getenv(name); return v != nullptr && v[0] != '\0' && v[0] != '0'; }

extern "C" int cantor_vecenv_create(cantor_vecenv** out, const cantor_env_params* params, int32_t precision,
                                    int64_t n_envs, int32_t device, int32_t n_chunks) {
    CANTOR_REQUIRE(out != nullptr && params != nullptr, "out/params is NULL");
    CANTOR_REQUIRE(precision == CANTOR_F32 || precision == CANTOR_F64, "precision must be 32 or 64");
    CANTOR_REQUIRE(n_envs > 0, "n_envs must be positive");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return fail(CANTOR_ERR_NO_DEVICE, "%s: %s", "cantor_vecenv_create", "no usable CUDA device (there is no CPU fallback)");
    cantor_vecenv* e = new (std::nothrow) cantor_vecenv();
    CANTOR_REQUIRE(e != nullptr, "out of host memory");
    e->params = *params;
    e->precision = precision;
    e->device = device;
    e->n_envs = n_envs;
    e->direct_small_outputs = !env_flag("CANTOR_HOST_NO_DIRECT");
    if (n_chunks <= 0) n_chunks = n_envs >= (1 << 16) ? 8 : 1;
    e->n_chunks = (int)(n_chunks > n_envs ? n_envs : n_chunks);
    e->rule = cantor_reset_rule{CANTOR_RESET_SAME_PATH, 0, nullptr, 0x5EED5EEDull, 0, 0};
    cudaError_t err = cudaSetDevice(device);
    for (int i = 0; i < 3 && err == cudaSuccess; ++i) err = cudaStreamCreateWithFlags(&e->streams[i], cudaStreamNonBlocking);
    const size_t rb = reward_bytes(e);
    if (err == cudaSuccess) err = cudaMalloc(&e->core, n_envs * 16);
    if (err == cudaSuccess) err = cudaMalloc(&e->cash, n_envs * rb);
    if (err == cudaSuccess && precision == CANTOR_F64) err = cudaMalloc(&e->pv_prev, n_envs * sizeof(double));
    if (err == cudaSuccess) err = cudaMalloc(&e->actions, n_envs * 2 * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&e->obs, n_envs * CANTOR_OBS_DIM * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&e->reward, n_envs * rb);
    if (err == cudaSuccess) err = cudaMalloc(&e->done, n_envs);
    if (err == cudaSuccess) err = cudaMalloc(&e->next_path, n_envs * sizeof(int32_t));
    if (err != cudaSuccess) {
        free_all(e);
        return cuda_fail(err, "cantor_vecenv_create");
    }
    *out = e;
    return CANTOR_OK;
}

extern "C" int cantor_vecenv_destroy(cantor_vecenv* env) {
    if (env == nullptr) return CANTOR_OK;
    return free_all(env);
}

extern "C" int cantor_vecenv_load_book_host(cantor_vecenv* e, const void* paths, const void* vols, const void* calls,
                                            const void* puts, int32_t src_dtype, int32_t n_paths, int32_t episode_length) {
    CANTOR_REQUIRE(e && paths && vols && calls && puts, "NULL argument");
    CANTOR_REQUIRE(src_dtype == CANTOR_F32 || src_dtype == CANTOR_F64, "src_dtype must be 32 or 64");
    int rc = alloc_book(e, n_paths, episode_length);
    if (rc) return rc;
    const size_t es = src_dtype == CANTOR_F64 ? 8 : 4;
    const size_t n1 = (size_t)n_paths * (episode_length + 1) * es, n0 = (size_t)n_paths * episode_length * es;
    void *dp = nullptr, *dv = nullptr, *dc = nullptr, *dq = nullptr;
    cudaStream_t s = e->streams[0];
    cudaError_t err = cudaMalloc(&dp, n1);
    if (err == cudaSuccess) err = cudaMalloc(&dv, n1);
    if (err == cudaSuccess) err = cudaMalloc(&dc, n0);
    if (err == cudaSuccess) err = cudaMalloc(&dq, n0);
    if (err == cudaSuccess) err = cudaMemcpyAsync(dp, paths, n1, cudaMemcpyHostToDevice, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(dv, vols, n1, cudaMemcpyHostToDevice, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(dc, calls, n0, cudaMemcpyHostToDevice, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(dq, puts, n0, cudaMemcpyHostToDevice, s);
    rc = CANTOR_OK;
    if (err == cudaSuccess) rc = cantor_pack_book(dp, dv, dc, dq, src_dtype, n_paths, episode_length, e->book, e->ld, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    cudaFree(dp); cudaFree(dv); cudaFree(dc); cudaFree(dq);
    if (err != cudaSuccess) return cuda_fail(err, "cantor_vecenv_load_book_host");
    return rc;
}

extern "C" int cantor_vecenv_simulate_book(cantor_vecenv* e, const cantor_sim_params* sim, int32_t n_paths,
                                           int32_t episode_length) {
    CANTOR_REQUIRE(e && sim, "NULL argument");
    int rc = alloc_book(e, n_paths, episode_length);
    if (rc) return rc;
    rc = cantor_sim_paths(sim, n_paths, episode_length, e->book, e->ld, e->streams[0]);
    if (rc) return rc;
    CANTOR_CUDA(cudaStreamSynchronize(e->streams[0]));
    return CANTOR_OK;
}

extern "C" int cantor_vecenv_set_reset_rule(cantor_vecenv* e, int32_t mode, uint64_t seed, int64_t env_offset) {
    CANTOR_REQUIRE(e != nullptr, "env is NULL");
    CANTOR_REQUIRE(mode >= CANTOR_RESET_SAME_PATH && mode <= CANTOR_RESET_PHILOX, "reset mode");
    e->rule.mode = mode;
    e->rule.seed = seed;
    e->rule.env_offset = env_offset;
    return CANTOR_OK;
}

extern "C" int cantor_vecenv_reset_host(cantor_vecenv* e, const int32_t* path_idx, float* obs_host) {
    CANTOR_REQUIRE(e != nullptr && e->book != nullptr, "env has no book (load or simulate one first)");
    CANTOR_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = e->streams[0];
    std::vector<int32_t> seq;
    if (path_idx == nullptr) {                       // env i starts on path (env_offset + i) mod n_paths
        seq.resize(e->n_envs);
        for (int64_t i = 0; i < e->n_envs; ++i) seq[i] = (int32_t)((e->rule.env_offset + i) % e->n_paths);
        path_idx = seq.data();
    } else {
        for (int64_t i = 0; i < e->n_envs; ++i)
            if (path_idx[i] < 0 || path_idx[i] >= e->n_paths) return fail(CANTOR_ERR_INVALID, "%s: %s", "cantor_vecenv_reset_host", "path index out of range");
    }
    CANTOR_CUDA(cudaMemcpyAsync(e->next_path, path_idx, e->n_envs * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    const cantor_replay_book b = book_of(e);
    const cantor_env_state st = state_of(e, 0);
    int rc = cantor_env_reset(&e->params, &b, &st, e->n_envs, e->precision, nullptr, e->next_path, e->obs, s);
    if (rc) return rc;
    if (obs_host) CANTOR_CUDA(cudaMemcpyAsync(obs_host, e->obs, e->n_envs * CANTOR_OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, s));
    CANTOR_CUDA(cudaStreamSynchronize(s));
    e->global_step = 0;
    return CANTOR_OK;
}

extern "C" int cantor_vecenv_step_host(cantor_vecenv* e, const float* actions_host, float* obs_host, void* reward_host,
                                       uint8_t* done_host, const int32_t* next_path_host) {
    CANTOR_REQUIRE(e != nullptr && e->book != nullptr, "env has no book (load or simulate one first)");
    CANTOR_REQUIRE(actions_host && obs_host && reward_host && done_host, "host buffer is NULL");
    CANTOR_CUDA(cudaSetDevice(e->device));
    const cantor_replay_book b = book_of(e);
    const size_t rb = reward_bytes(e);
    cantor_reset_rule rule = e->rule;
    rule.episode_counter = e->global_step;
    if (next_path_host != nullptr) {
        CANTOR_CUDA(cudaMemcpyAsync(e->next_path, next_path_host, e->n_envs * sizeof(int32_t), cudaMemcpyHostToDevice, e->streams[0]));
        CANTOR_CUDA(cudaStreamSynchronize(e->streams[0]));
        rule.mode = CANTOR_RESET_FROM_ARRAY;
    }
    // Page-locked, device-mapped result buffers (cantor_host_register / cudaHostAlloc): the step kernel writes the 5 small
    // bytes per env -- reward and done -- straight into them over PCIe, which saves two of the three D2H copies of every chunk
    // (16 copy-engine launches per step at 8 chunks: 1.264 -> 1.229 ms per 2^20-env step); the observations (52 B per env) still go
    // by DMA -- letting the kernel's TMA stores write them into host memory too measured slower (1.259 ms).
    auto mapped = [](const void* host) -> void* {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (at.type == cudaMemoryTypeHost && at.devicePointer != nullptr) ? at.devicePointer : nullptr;
    };
    char* reward_direct = (char*)mapped(reward_host);
    uint8_t* done_direct = (uint8_t*)mapped(done_host);
    const bool direct = e->direct_small_outputs && reward_direct != nullptr && done_direct != nullptr;
    const int C = e->n_chunks;
    // chunk boundaries on multiples of 128 envs so every chunk keeps the aligned TMA obs store
    const int64_t per = ((e->n_envs + C - 1) / C + 127) / 128 * 128;
    for (int c = 0; c < C; ++c) {
        const int64_t first = (int64_t)c * per;
        if (first >= e->n_envs) break;
        const int64_t n = (first + per <= e->n_envs) ? per : e->n_envs - first;
        cudaStream_t s = e->streams[c % 3];
        CANTOR_CUDA(cudaMemcpyAsync(e->actions + first * 2, actions_host + first * 2, n * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
        const cantor_env_state st = state_of(e, first);
        cantor_reset_rule r = rule;
        r.env_offset = rule.env_offset + first;
        r.next_path = e->next_path + first;
        int rc = cantor_env_step(&e->params, &b, &st, n, e->precision, e->actions + first * 2,
                                 e->obs + first * CANTOR_OBS_DIM, direct ? (void*)(reward_direct + first * rb) : (void*)((char*)e->reward + first * rb),
                                 direct ? done_direct + first : e->done + first, nullptr, 1, &r, nullptr, s);
        if (rc) return rc;
        CANTOR_CUDA(cudaMemcpyAsync(obs_host + first * CANTOR_OBS_DIM, e->obs + first * CANTOR_OBS_DIM,
                                    n * CANTOR_OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (!direct) {
            CANTOR_CUDA(cudaMemcpyAsync((char*)reward_host + first * rb, (char*)e->reward + first * rb, n * rb, cudaMemcpyDeviceToHost, s));
            CANTOR_CUDA(cudaMemcpyAsync(done_host + first, e->done + first, n, cudaMemcpyDeviceToHost, s));
        }
    }
    for (auto& s : e->streams) CANTOR_CUDA(cudaStreamSynchronize(s));
    e->global_step += 1;
    return CANTOR_OK;
}

extern "C" int cantor_vecenv_episode_length(const cantor_vecenv* e) { return e ? e->T : -1; }
extern "C" int cantor_vecenv_num_paths(const cantor_vecenv* e) { return e ? e->n_paths : -1; }

// Page-lock / unlock a caller-owned host buffer so the per-step copies run at full PCIe speed and asynchronously.
extern "C" int cantor_host_register(void* ptr, size_t bytes) {
    CANTOR_REQUIRE(ptr != nullptr && bytes > 0, "NULL / empty buffer");
    CANTOR_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
    return CANTOR_OK;
}
extern "C" int cantor_host_unregister(void* ptr) {
    CANTOR_REQUIRE(ptr != nullptr, "NULL buffer");
    CANTOR_CUDA(cudaHostUnregister(ptr));
    return CANTOR_OK;
}

// Raw ceiling of the host-buffer face: the copies of cantor_vecenv_step_host and nothing else.  One round moves h2d_bytes
// host -> device and d2h_bytes device -> host between page-locked host memory and HBM in n_chunks pieces round-robined over
// three streams (the H2D piece of a chunk, then its D2H piece, on the chunk's stream), then -- with sync_each_round, as a gym
// step must -- waits for all three streams.  Repeats for at least `seconds` of wall time.
extern "C" int cantor_host_copy_probe(int32_t device, int64_t d2h_bytes, int64_t h2d_bytes, int32_t n_chunks,
                                      int32_t sync_each_round, double seconds, double* d2h_gbs, double* h2d_gbs,
                                      double* rounds_per_s) {
    CANTOR_REQUIRE(d2h_bytes >= 0 && h2d_bytes >= 0 && d2h_bytes + h2d_bytes > 0, "nothing to copy");
    CANTOR_REQUIRE(n_chunks >= 1 && n_chunks <= 4096 && seconds > 0 && seconds <= 60, "n_chunks in [1, 4096], seconds in (0, 60]");
    CANTOR_REQUIRE(d2h_gbs && h2d_gbs && rounds_per_s, "output pointer is NULL");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return fail(CANTOR_ERR_NO_DEVICE, "%s: %s", "cantor_host_copy_probe", "no usable CUDA device");
    CANTOR_CUDA(cudaSetDevice(device));
    void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    cudaStream_t st[3] = {nullptr, nullptr, nullptr};
    cudaError_t err = cudaSuccess;
    if (h2d_bytes) err = cudaHostAlloc(&h_in, h2d_bytes, cudaHostAllocPortable);
    if (err == cudaSuccess && d2h_bytes) err = cudaHostAlloc(&h_out, d2h_bytes, cudaHostAllocPortable);
    if (err == cudaSuccess && h2d_bytes) err = cudaMalloc(&d_in, h2d_bytes);
    if (err == cudaSuccess && d2h_bytes) err = cudaMalloc(&d_out, d2h_bytes);
    if (err == cudaSuccess && h_in) memset(h_in, 1, h2d_bytes);
    if (err == cudaSuccess && h_out) memset(h_out, 0, d2h_bytes);
    for (int i = 0; i < 3 && err == cudaSuccess; ++i) err = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    long long rounds = 0;
    double elapsed = 0.0;
    if (err == cudaSuccess) {
        const int64_t din = (h2d_bytes / n_chunks + 15) / 16 * 16, dout = (d2h_bytes / n_chunks + 15) / 16 * 16;
        auto one_round = [&]() -> cudaError_t {
            cudaError_t e = cudaSuccess;
            for (int c = 0; c < n_chunks && e == cudaSuccess; ++c) {
                const int64_t i0 = (int64_t)c * din, o0 = (int64_t)c * dout;
                const int64_t ni = i0 < h2d_bytes ? (i0 + din <= h2d_bytes ? din : h2d_bytes - i0) : 0;
                const int64_t no = o0 < d2h_bytes ? (o0 + dout <= d2h_bytes ? dout : d2h_bytes - o0) : 0;
                if (ni > 0) e = cudaMemcpyAsync((char*)d_in + i0, (char*)h_in + i0, ni, cudaMemcpyHostToDevice, st[c % 3]);
                if (e == cudaSuccess && no > 0) e = cudaMemcpyAsync((char*)h_out + o0, (char*)d_out + o0, no, cudaMemcpyDeviceToHost, st[c % 3]);
            }
            return e;
        };
        auto sync_all = [&]() -> cudaError_t {
            cudaError_t e = cudaSuccess;
            for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaStreamSynchronize(st[i]);
            return e;
        };
        err = one_round();                                   // warm-up
        if (err == cudaSuccess) err = sync_all();
        const auto t0 = std::chrono::steady_clock::now();
        while (err == cudaSuccess) {
            err = one_round();
            if (err == cudaSuccess && (sync_each_round || (rounds & 7) == 7)) err = sync_all();
            ++rounds;
            elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (elapsed >= seconds && (sync_each_round || (rounds & 7) == 0)) break;
        }
        if (err == cudaSuccess) err = sync_all();
        elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    for (auto& s : st) if (s) cudaStreamDestroy(s);
    cudaFreeHost(h_in); cudaFreeHost(h_out); cudaFree(d_in); cudaFree(d_out);
    if (err != cudaSuccess) return cuda_fail(err, "cantor_host_copy_probe");
    *rounds_per_s = rounds / elapsed;
    *d2h_gbs = (double)d2h_bytes * rounds / elapsed / 1e9;
    *h2d_gbs = (double)h2d_bytes * rounds / elapsed / 1e9;
    return CANTOR_OK;
}
