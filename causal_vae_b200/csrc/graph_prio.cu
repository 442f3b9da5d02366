// Per-node launch priorities for a captured training step.
//
// The step is captured with its weight-gradient family forked onto a side stream (chain.py::side_wgrad).  Inside a
// replayed graph both branches compete for SMs at CTA granularity, and a stream-ordered replay gives every node the
// priority of the LAUNCH stream: whichever kernel was released first owns the SMs until its grid has drained.  The
// chain of input gradients + BatchNorm-backward passes is the critical path, yet its small grids queued behind
// multi-wave weight-gradient grids of the side branch (timeline: 0.7 ms of such waits per step,
// profiles/r2_step_timeline.txt).  cudaGraphInstantiateFlagUseNodePriority makes the replay honour per-node
// priorities instead; this file sets them (by kernel family) and owns the executable graph.
#include <cstring>
#include "common.cuh"

namespace {

bool side_family(const char* name) {
  // weight gradients, their split-K reductions and the bias column sums: never on the critical path
  return std::strstr(name, "wgrad") != nullptr || std::strstr(name, "col_reduce") != nullptr;
}

}  // namespace

// Sets cudaLaunchAttributePriority on every kernel node of `graph`: `prio_main` for the main chain, `prio_side` for the
// weight-gradient family (CUDA convention: lower number = higher priority).  counts[0] = kernel nodes, counts[1] = nodes
// given prio_side, counts[2] = nodes whose function name could not be resolved (they get prio_main).
extern "C" int cvae_graph_set_priorities(void* graph, int prio_main, int prio_side, int* counts) {
  if (!graph) return CVAE_ERR_BAD_ARG;
  cudaGraph_t g = static_cast<cudaGraph_t>(graph);
  size_t n = 0;
  if (cudaGraphGetNodes(g, nullptr, &n) != cudaSuccess) return CVAE_ERR_LAUNCH;
  cudaGraphNode_t* nodes = new cudaGraphNode_t[n ? n : 1];
  int rc = CVAE_OK, nk = 0, ns = 0, nu = 0;
  if (cudaGraphGetNodes(g, nodes, &n) != cudaSuccess) rc = CVAE_ERR_LAUNCH;
  for (size_t i = 0; rc == CVAE_OK && i < n; ++i) {
    cudaGraphNodeType t;
    if (cudaGraphNodeGetType(nodes[i], &t) != cudaSuccess) { rc = CVAE_ERR_LAUNCH; break; }
    if (t != cudaGraphNodeTypeKernel) continue;
    ++nk;
    bool side = false;
    cudaKernelNodeParams kp;
    const char* name = nullptr;
    if (cudaGraphKernelNodeGetParams(nodes[i], &kp) == cudaSuccess && kp.func != nullptr &&
        cudaFuncGetName(&name, kp.func) == cudaSuccess && name != nullptr) {
      side = side_family(name);
    } else {
      (void)cudaGetLastError();
      ++nu;
    }
    ns += side ? 1 : 0;
    cudaLaunchAttributeValue v;
    std::memset(&v, 0, sizeof(v));
    v.priority = side ? prio_side : prio_main;
    if (cudaGraphKernelNodeSetAttribute(nodes[i], cudaLaunchAttributePriority, &v) != cudaSuccess) { rc = CVAE_ERR_LAUNCH; break; }
  }
  delete[] nodes;
  if (counts) { counts[0] = nk; counts[1] = ns; counts[2] = nu; }
  if (rc != CVAE_OK) (void)cudaGetLastError();
  return rc;
}

// cudaGraphInstantiateWithFlags(UseNodePriority).  The caller keeps `graph` (and the memory pool it was captured in) alive
// for as long as the executable graph is launched.
extern "C" int cvae_graph_instantiate_prio(void* graph, void** exec_out) {
  if (!graph || !exec_out) return CVAE_ERR_BAD_ARG;
  cudaGraphExec_t e = nullptr;
  if (cudaGraphInstantiateWithFlags(&e, static_cast<cudaGraph_t>(graph), cudaGraphInstantiateFlagUseNodePriority) != cudaSuccess) {
    (void)cudaGetLastError();
    return CVAE_ERR_LAUNCH;
  }
  *exec_out = e;
  return CVAE_OK;
}

extern "C" int cvae_graph_launch(void* exec, cvae_stream_t s) {
  if (!exec) return CVAE_ERR_BAD_ARG;
  return cudaGraphLaunch(static_cast<cudaGraphExec_t>(exec), as_stream(s)) == cudaSuccess ? CVAE_OK : CVAE_ERR_LAUNCH;
}

extern "C" int cvae_graph_exec_destroy(void* exec) {
  if (!exec) return CVAE_OK;
  return cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(exec)) == cudaSuccess ? CVAE_OK : CVAE_ERR_LAUNCH;
}
