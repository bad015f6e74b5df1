/*
 * multi_node.c -- BASELINE config #1's shape (the reference's examples/multi-node: ONE Raft
 * group, THREE nodes) driven from plain C through the C ABI only.  No torch, no CUDA headers:
 * this is what a cgo / Rust-FFI / JNI binding sees.
 *
 *   gcc -I include examples/multi_node.c josefine_b200/csrc/libjosefine_b200.so -o multi_node
 *
 * It ticks the three co-resident nodes every 100 ms of logical time (server.rs:25) from a cold
 * start, waits for the seeded election timeouts to produce a leader, proposes three client
 * requests to it (server.rs:156-160), and checks that every node applies them in order.
 * The leader applies (prev..=new).skip(1) (leader.rs:93); a follower applies the HALF-OPEN
 * range prev..commit (follower.rs:204), i.e. it trails by one block until the next commit --
 * reference behaviour, reproduced.  Prints one line per event; exit code 0 on success.
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

int main(void) {
  jr_config cfg;
  jr_config_default(&cfg, 1, 3);
  cfg.seed = 42;
  cfg.flags = JR_F_CAPTURE_MESSAGES | JR_F_CAPTURE_FSM;
  jr_engine* e = NULL;
  CHECK(jr_engine_create(&cfg, &e));

  jr_msg msgs[256];
  jr_fsm_instr fsm[256];
  jr_proposal prop[1];
  jr_step_args a;
  uint64_t now = 0;
  uint32_t leader = 0;
  int proposed = 0, applied[4] = {0, 0, 0, 0};
  const uint64_t tokens[3] = {0xA1, 0xB2, 0xC3};

  for (int tick = 1; tick <= 60; ++tick) {
    now += 100;
    memset(&a, 0, sizeof a);
    a.now_ms = now;
    a.flags = JR_STEP_DELIVER | JR_STEP_TICK;
    a.out_msgs = msgs; a.cap_msgs = 256;
    a.out_fsm = fsm;  a.cap_fsm = 256;
    if (leader && proposed < 3 && tick % 2 == 0) {   /* a client request, every other tick */
      prop[0].token = tokens[proposed];
      prop[0].node = leader;
      prop[0].reserved = 0;
      a.proposals = prop;
      printf("tick %2d: propose token %#llx to node %u\n", tick, (unsigned long long)tokens[proposed], leader);
      ++proposed;
    }
    CHECK(jr_step(e, &a));
    for (size_t i = 0; i < a.n_fsm; ++i)
      if (fsm[i].kind == JR_FSM_APPLY && fsm[i].block.id != 0) {  /* fsm.rs:61-63 skips block 0 */
        int k = applied[fsm[i].node];
        if (k >= 3 || fsm[i].block.data != tokens[k]) {
          fprintf(stderr, "node %u applied token %#llx out of order\n", fsm[i].node, (unsigned long long)fsm[i].block.data);
          return 1;
        }
        applied[fsm[i].node] = k + 1;
        printf("tick %2d: node %u applies block %llu (token %#llx)\n", tick, fsm[i].node,
               (unsigned long long)fsm[i].block.id, (unsigned long long)fsm[i].block.data);
      }
    if (!leader) {
      jr_leader_entry le;
      CHECK(jr_leader_table(e, &le));
      if (le.leader_id) {
        leader = le.leader_id;
        printf("tick %2d: node %u is leader of term %llu\n", tick, leader, (unsigned long long)le.term);
      }
    }
  }
  int ok = leader != 0;
  for (uint32_t n = 1; n <= 3; ++n) {
    jr_replica_state st;
    CHECK(jr_query(e, 0, n, &st));
    printf("node %u: role %u term %llu head %llu commit %llu fault %u applied %d\n", n, st.role,
           (unsigned long long)st.current_term, (unsigned long long)st.head, (unsigned long long)st.commit, st.fault,
           applied[n]);
    ok = ok && st.fault == 0 && st.commit == 3 && applied[n] == (n == leader ? 3 : 2);
  }
  jr_engine_destroy(e);
  printf(ok ? "OK\n" : "FAILED\n");
  return ok ? 0 : 1;
}
