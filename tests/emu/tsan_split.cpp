// TEST INFRASTRUCTURE ONLY -- race check of the split-launch hand-over.
//
// Built with -fsanitize=thread together with the engine's device code compiled for the host
// (cuda_emu.h, JR_EMU_CTAS CTAs at a time).  A launch cut into JR_PARTS tasks hands a block's
// state and mailboxes from one CTA to another through global memory; the only ordering is the
// release store / acquire load on done[block] plus the CTA barriers.  ThreadSanitizer sees
// every plain access of the emulated device code, so a missing edge shows up as a data race.
// With -DJR_EMU_BREAK_HANDOFF (relaxed flag accesses) it must report one: the negative control.
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <vector>

#include "../../include/josefine_raft_abi.h"

static jr_engine* make(const char* parts, uint32_t G, uint32_t R) {
  setenv("JR_PARTS", parts, 1);
  jr_config cfg;
  jr_config_default(&cfg, G, R);
  cfg.seed = 3;
  cfg.flags = JR_F_STREAM_DIGEST;
  cfg.chain_capacity = 256;
  jr_engine* e = nullptr;
  if (jr_engine_create(&cfg, &e) != JR_OK) { fprintf(stderr, "create: %s\n", jr_last_error()); exit(2); }
  return e;
}

int main() {
  const uint32_t G = getenv("JR_TSAN_GROUPS") ? (uint32_t)atoi(getenv("JR_TSAN_GROUPS")) : 32, R = 3;  // one 32-group block: its parts overlap pairwise at JR_EMU_CTAS=2 (TSAN tracks <= 256 live threads)
  jr_engine* split = make("4", G, R);
  jr_engine* whole = make("1", G, R);
  for (jr_engine* e : {split, whole}) {
    if (jr_run(e, 100, 100, 22, 0) != JR_OK || jr_run(e, 2300, 100, 21, 2) != JR_OK) {
      fprintf(stderr, "run: %s\n", jr_last_error());
      return 2;
    }
  }
  uint64_t ds = 0, dw = 0, ms, fs, ns, nfs, mw, fw, nw, nfw, faults = 0;
  jr_state_digest(split, &ds);
  jr_state_digest(whole, &dw);
  jr_stream_digest(split, &ms, &fs, &ns, &nfs);
  jr_stream_digest(whole, &mw, &fw, &nw, &nfw);
  jr_fault_count(split, &faults);
  std::vector<jr_leader_entry> tabv(G);
  jr_leader_entry* tab = tabv.data();
  jr_leader_table(split, tab);
  unsigned leaders = 0, committed = 0;
  for (uint32_t g = 0; g < G; ++g) { leaders += tab[g].leader_id != 0; committed += tab[g].commit > 0; }
  jr_engine_destroy(split);
  jr_engine_destroy(whole);
  printf("state %llx/%llx msgs %llu/%llu fsm %llu/%llu leaders %u committed %u faults %llu\n", (unsigned long long)ds,
         (unsigned long long)dw, (unsigned long long)ns, (unsigned long long)nw, (unsigned long long)nfs,
         (unsigned long long)nfw, leaders, committed, (unsigned long long)faults);
  if (ds != dw || ms != mw || fs != fw || ns != nw || nfs != nfw) { printf("MISMATCH\n"); return 1; }
  if (leaders < G / 2 || committed == 0) { printf("workload did not get going\n"); return 1; }
  printf("split == whole\n");
  return 0;
}
