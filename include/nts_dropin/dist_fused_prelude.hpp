// Forced include (-include) for building the reference's host code with the device-resident distributed aggregation:
// with -DNTSDISTCPUFUSEDGRAPHOP_HPP the original core/ntsDistGPUFusedGraphOp.hpp is skipped by its own include guard,
// and this prelude supplies the same class right after the reference's umbrella header.  Nothing in the reference tree
// is modified.
#ifndef NTS_B200_DIST_FUSED_PRELUDE_HPP
#define NTS_B200_DIST_FUSED_PRELUDE_HPP
#ifndef NTSDISTCPUFUSEDGRAPHOP_HPP
#error "compile with -DNTSDISTCPUFUSEDGRAPHOP_HPP so that the original core/ntsDistGPUFusedGraphOp.hpp is skipped"
#endif
#include "core/neutronstar.hpp"
#include "nts_dropin/core/ntsDistGPUFusedGraphOp.hpp"
#endif
