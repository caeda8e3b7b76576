// TEST INFRASTRUCTURE ONLY (oracle build): single-node stand-in for libnuma.
// The reference CPU engine (core/graph.hpp:346-411 and friends) calls eleven
// libnuma entry points; this image has no libnuma, so the oracle build of the
// unmodified reference links against these trivial versions instead.  One NUMA
// node, `NTS_THREADS` (or the online CPU count) CPUs, plain heap memory.
#pragma once
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

struct bitmask {
  int unused;
};

static inline int numa_available() { return 0; }
static inline int numa_num_configured_nodes() { return 1; }
static inline int numa_num_configured_cpus() {
  const char *env = getenv("NTS_THREADS");
  if (env && atoi(env) > 0)
    return atoi(env);
  return (int)sysconf(_SC_NPROCESSORS_ONLN);
}
static inline void *numa_alloc_onnode(size_t bytes, int /*node*/) {
  return calloc(1, bytes ? bytes : 1);
}
static inline void *numa_alloc_interleaved(size_t bytes) {
  return calloc(1, bytes ? bytes : 1);
}
static inline void numa_free(void *p, size_t /*bytes*/) { free(p); }
static inline void *numa_realloc(void *p, size_t /*old_bytes*/, size_t new_bytes) {
  return realloc(p, new_bytes ? new_bytes : 1);
}
static inline void numa_tonode_memory(void *, size_t, int) {}
static inline struct bitmask *numa_parse_nodestring(const char *) {
  static struct bitmask all;
  return &all;
}
static inline void numa_set_interleave_mask(struct bitmask *) {}
static inline int numa_run_on_node(int) { return 0; }
