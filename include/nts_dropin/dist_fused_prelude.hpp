// Force-included (-include) ahead of toolkits/main.cpp by oracle/Makefile's `dropin_dist` target.  With
// -DNTSDISTCPUFUSEDGRAPHOP_HPP (the include guard [sic] of the original core/ntsDistGPUFusedGraphOp.hpp) the
// reference's own ForwardGPUfuseOp is skipped; its DistGPUGetDepNbrOp (core/ntsDistGPUGraphOp.hpp:48-143) is compiled
// under another name; both classes are then supplied by include/nts_dropin/core/ on top of the peer-memory exchange
// engine.  Nothing of the reference tree is modified or copied.
#pragma once
#ifndef NTSDISTCPUFUSEDGRAPHOP_HPP
#error "compile with -DNTSDISTCPUFUSEDGRAPHOP_HPP so that the original core/ntsDistGPUFusedGraphOp.hpp is skipped"
#endif
#define DistGPUGetDepNbrOp DistGPUGetDepNbrOp_host_staged
#include "core/neutronstar.hpp"
#undef DistGPUGetDepNbrOp
#include "nts_dropin/core/ntsDistGPUFusedGraphOp.hpp"
#include "nts_dropin/core/ntsDistGPUGetDepNbrOp.hpp"
#include "nts_dropin/core/ntsDistGPUFusedGATOp.hpp"
