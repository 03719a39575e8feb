// Host-side profile of the self-play loop without a GPU: links e55_game.cpp against stub network entry points that
// return cheap pseudo-random logits, so tree / move generation / mate search / encoding costs can be measured here.
//   g++ -O3 -std=c++20 -march=x86-64-v3 -pthread -I. tools/host_prof.cpp erweitern_55_b200/csrc/e55_game.cpp -o /tmp/host_prof
// The same build with -O1 -g -fsanitize=address,undefined or -fsanitize=thread is how the engine, the tree and the
// pooled self-play driver were checked for memory errors and data races (clean on 20 000 moves, 6 threads).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../include/e55.h"

static std::atomic<long> g_calls{0}, g_positions{0};

extern "C" {
static float g_table[8192];
static void fill(const uint32_t* bits, int n, float* policy, float* value) {
  if (g_table[1] == 0.f) {
    uint32_t x = 99u;
    for (float& v : g_table) { x = x * 1664525u + 1013904223u; v = static_cast<float>(x >> 8) * (4.f / 16777216.f) - 2.f; }
  }
  for (int i = 0; i < n; ++i) {   // about what the service's copy-out costs: one 6.9 KB memcpy per position
    const uint32_t x = (bits[i * 266] * 2654435761u + bits[i * 266 + 5] * 40503u + bits[i * 266 + 38] * 977u) >> 7;
    memcpy(policy + static_cast<size_t>(i) * 1725, g_table + x % (8192 - 1725), 1725 * sizeof(float));
    value[i] = g_table[x % 8192] * 0.5f;
  }
}
// what the service costs the caller: copy-in of the packed input and the move lists, copy-out of the priors
int e55_service_predict_packed_moves(e55_ctx*, const uint32_t* bits, const float* scale, int n, const int32_t* off,
                                     const uint16_t* idx, float* priors, float* value) {
  g_calls++; g_positions += n;
  static thread_local uint32_t stage_b[266 * 256];
  static thread_local float stage_s[266 * 256];
  static thread_local uint16_t stage_i[256 * 256];
  memcpy(stage_b, bits, sizeof(uint32_t) * 266 * n); memcpy(stage_s, scale, sizeof(float) * 266 * n);
  memcpy(stage_i, idx, sizeof(uint16_t) * off[n]);
  if (g_table[1] == 0.f) { float dummy[1725]; float v; fill(bits, 1, dummy, &v); }
  for (int i = 0; i < n; ++i) {
    const uint32_t x = (bits[i * 266] * 2654435761u + bits[i * 266 + 5] * 40503u + bits[i * 266 + 38] * 977u) >> 7;
    const int k = off[i + 1] - off[i];
    float sum = 0.f;
    for (int j = 0; j < k; ++j) { priors[off[i] + j] = 2.1f + g_table[(x + stage_i[off[i] + j]) % 8192]; sum += priors[off[i] + j]; }
    for (int j = 0; j < k; ++j) priors[off[i] + j] /= sum;
    value[i] = g_table[x % 8192] * 0.5f;
  }
  return 0;
}
int e55_predict_packed_moves(e55_ctx* c, const uint32_t* bits, const float* s, int n, const int32_t* off, const uint16_t* idx,
                             float* priors, float* value) {
  return e55_service_predict_packed_moves(c, bits, s, n, off, idx, priors, value);
}
// asynchronous form: answered at once, handed over at fetch
struct Pending { std::vector<float> priors, value; };
static std::mutex g_mu;
static std::vector<Pending*> g_pending;
int e55_service_submit_packed_moves(e55_ctx* c, const uint32_t* bits, const float* s, int n, const int32_t* off, const uint16_t* idx,
                                    int, e55_ticket* t) {
  Pending* p = new Pending{std::vector<float>(off[n]), std::vector<float>(n)};
  e55_service_predict_packed_moves(c, bits, s, n, off, idx, p->priors.data(), p->value.data());
  std::lock_guard<std::mutex> lk(g_mu);
  int slot = -1;
  for (size_t i = 0; i < g_pending.size(); ++i) if (!g_pending[i]) { slot = static_cast<int>(i); break; }
  if (slot < 0) { slot = static_cast<int>(g_pending.size()); g_pending.push_back(nullptr); }
  g_pending[slot] = p;
  t->batch = slot; t->n = n; t->total = off[n];
  return 0;
}
int e55_service_fetch(e55_ctx*, const e55_ticket* t, float* priors, float* value, int) {
  Pending* p;
  { std::lock_guard<std::mutex> lk(g_mu); p = g_pending[t->batch]; g_pending[t->batch] = nullptr; }
  memcpy(priors, p->priors.data(), sizeof(float) * p->priors.size());
  memcpy(value, p->value.data(), sizeof(float) * p->value.size());
  delete p;
  return 0;
}
int e55_get_arch(const e55_ctx*, int* b, int* f, int* ip) { *b = 7; *f = 128; *ip = 266; return 0; }
int e55_service_stats(e55_ctx*, int64_t* b, int64_t* p) { *b = g_calls; *p = g_positions; return 0; }
}

int main(int argc, char** argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 1;
  const long moves = argc > 2 ? atol(argv[2]) : 200;
  const int mate = argc > 3 ? atoi(argv[3]) : 7;
  const int sims = argc > 4 ? atoi(argv[4]) : 800;
  { uint32_t b[266] = {}; float pol[1725], v; fill(b, 1, pol, &v); }   // the stub's logit table, before any worker thread looks at it
  e55_selfplay_config cfg;
  e55_selfplay_default_config(&cfg);
  cfg.threads = threads; cfg.total_moves = moves; cfg.mate_depth = mate; cfg.simulations = sims;
  cfg.games_per_thread = argc > 5 ? atoi(argv[5]) : 1;
  e55_selfplay_stats st;
  const int rc = e55_selfplay_run_ex(reinterpret_cast<e55_ctx*>(0x1), &cfg, &st);
  printf("rc %d: %ld moves, %ld evals, %ld games (%.1f plies) in %.3f s -> %.1f moves/s, %.2f us/eval, %.2f ms/move/thread\n", rc,
         st.moves, st.evals, st.games, st.mean_game_plies, st.seconds, st.moves / st.seconds,
         1e6 * st.seconds * threads / (st.evals > 0 ? st.evals : 1), 1e3 * st.seconds * threads / st.moves);
  return rc;
}
