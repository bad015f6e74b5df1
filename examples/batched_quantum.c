/*
 * batched_quantum.c -- the hot path from plain C: thousands of Raft groups stepped a quantum at a time through the
 * C ABI only (no torch, no CUDA headers -- what a cgo / Rust-FFI / JNI binding sees), the way INTEGRATION.md
 * section 2a describes and bench.py's end-to-end leg measures:
 *
 *   per quantum of TICKS ticks:  jr_run_token_runs   proposals in, run-length: {base, stride} per group
 *                                jr_leader_table_async   who leads, term, commit -> pinned host memory
 *                                jr_fsm_records_async    the quantum's Instruction stream, compact -> pinned host memory
 *   one quantum later:           jr_leader_table_wait, jr_fsm_records_wait, jr_fsm_fold (or jr_fsm_expand)
 *
 *   gcc -I include examples/batched_quantum.c josefine_b200/csrc/libjosefine_b200.so -o batched_quantum
 *   ./batched_quantum [groups] [quanta]
 *
 * Cold start: 3 replicas per group, seeded election timeouts (follower.rs:103-113), one election per group (with R = 3
 * a candidate wins whatever the delivery order of the duplicate VoteRequests, SURVEY note N3).  Then every quantum
 * proposes one block per group and tick to the announced leader.  Checks, per quantum: no record dropped, one Notify
 * per proposal, every replica's apply watermark at most 3 blocks behind the leader's commit (a follower hears the commit
 * with the next heartbeat -- every second tick at heartbeat_ms = tick, leader.rs:78-84 -- and applies the half-open range
 * prev..commit, follower.rs:204), no faulted replica.  Exit code 0 and a last line "OK" on success.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "josefine_raft_abi.h"

#define CHECK(call)                                                               \
  do {                                                                            \
    jr_status _s = (call);                                                        \
    if (_s != JR_OK) {                                                            \
      fprintf(stderr, "%s -> status %d: %s\n", #call, (int)_s, jr_last_error());  \
      return 2;                                                                   \
    }                                                                             \
  } while (0)

enum { R = 3, TICKS = 32, DT_MS = 100, DEPTH = 2 };   /* DEPTH quanta in flight (< JR_STAGING_DEPTH) */

int main(int argc, char** argv) {
  const uint32_t G = argc > 1 ? (uint32_t)atoi(argv[1]) : 2048u;
  const int quanta = argc > 2 ? atoi(argv[2]) : 8;
  jr_config cfg;
  jr_config_default(&cfg, G, R);
  cfg.seed = 7;
  cfg.flags = JR_F_CAPTURE_FSM;
  cfg.chain_capacity = 256;
  cfg.fsm_units = 64;
  jr_engine* e = NULL;
  CHECK(jr_engine_create(&cfg, &e));

  /* cold start -> one leader per group */
  uint64_t now = DT_MS;
  CHECK(jr_run(e, now, DT_MS, 100, 0));
  now += 100 * DT_MS;
  jr_leader_entry* table[DEPTH];
  jr_token_run* runs[DEPTH];
  for (int b = 0; b < DEPTH; ++b) {
    CHECK(jr_host_alloc((size_t)G * sizeof(jr_leader_entry), (void**)&table[b]));
    CHECK(jr_host_alloc((size_t)G * sizeof(jr_token_run), (void**)&runs[b]));
  }
  CHECK(jr_leader_table(e, table[0]));   /* first announce: where the proposals go */
  uint32_t led = 0;
  for (uint32_t g = 0; g < G; ++g) led += table[0][g].leader_id != 0;
  printf("%u of %u groups elected a leader within 100 ticks\n", led, G);
  if (led != G) { fprintf(stderr, "expected every group to elect a leader\n"); return 1; }
  { /* the start-up's Instructions (block 0 applies, ...) are not part of the quanta below */
    const jr_fsm_record* recs; jr_fsm_batch batch;
    CHECK(jr_fsm_records_async(e));
    CHECK(jr_fsm_records_wait(e, &recs, &batch));
  }
  CHECK(jr_set_auto_truncate(e, 1, 8));   /* D7: every fused run ends with jr_truncate(8) */

  uint32_t* applied = (uint32_t*)calloc((size_t)G * R, sizeof(uint32_t));
  uint64_t totals[3] = {0, 0, 0};          /* Apply instructions, Notify instructions, records */
  int submitted = 0, consumed = 0, rc = 0;
  while (consumed < quanta) {
    while (submitted < quanta && submitted - consumed < DEPTH) {   /* submit: asynchronous */
      const int b = submitted % DEPTH;
      for (uint32_t g = 0; g < G; ++g) {
        runs[b][g].base = ((uint64_t)(submitted * TICKS + 1) << 32) | (g + 1);   /* request numbers: never 0, never repeated */
        runs[b][g].stride = 1ull << 32;
      }
      CHECK(jr_run_token_runs(e, now, DT_MS, TICKS, runs[b]));
      now += (uint64_t)TICKS * DT_MS;
      CHECK(jr_leader_table_async(e, table[b]));
      CHECK(jr_fsm_records_async(e));
      ++submitted;
    }
    /* consume the oldest quantum while the next one runs */
    const int b = consumed % DEPTH;
    const jr_fsm_record* recs;
    jr_fsm_batch batch;
    CHECK(jr_leader_table_wait(e));
    CHECK(jr_fsm_records_wait(e, &recs, &batch));
    const uint64_t notify_before = totals[1];
    CHECK(jr_fsm_fold(recs, batch.n_records, G, R, applied, totals));
    const uint64_t notifies = totals[1] - notify_before;
    uint32_t lag_max = 0;
    for (uint32_t g = 0; g < G; ++g)
      for (uint32_t n = 0; n < R; ++n) {
        const uint32_t commit = table[b][g].commit, hi = applied[(size_t)n * G + g];
        const uint32_t lag = commit > hi ? commit - hi : 0;
        if (lag > lag_max) lag_max = lag;
      }
    printf("quantum %d: %llu records for %llu Instructions (%llu Notify), commit of group 0 = %u, apply lag <= %u\n", consumed,
           (unsigned long long)batch.n_records, (unsigned long long)batch.n_instructions, (unsigned long long)notifies,
           table[b][0].commit, lag_max);
    if (batch.n_dropped || notifies != (uint64_t)G * TICKS || lag_max > 3) {
      fprintf(stderr, "quantum %d: dropped %llu, notifies %llu (expected %llu), lag %u\n", consumed,
              (unsigned long long)batch.n_dropped, (unsigned long long)notifies, (unsigned long long)G * TICKS, lag_max);
      rc = 1;
    }
    ++consumed;
  }
  uint64_t faults = 0;
  CHECK(jr_fault_count(e, &faults));
  if (faults) { fprintf(stderr, "%llu replicas faulted\n", (unsigned long long)faults); rc = 1; }
  printf("%llu Apply + %llu Notify instructions in %llu records over %d quanta of %d ticks\n", (unsigned long long)totals[0],
         (unsigned long long)totals[1], (unsigned long long)totals[2], quanta, TICKS);
  for (int b = 0; b < DEPTH; ++b) { jr_host_free(table[b]); jr_host_free(runs[b]); }
  free(applied);
  jr_engine_destroy(e);
  if (rc == 0) printf("OK\n");
  return rc;
}
