// dp.cuh - data-parallel plane of libinb200: one process per GPU, the batch sharded along its
// outermost dimension, parameters replicated.  Everything here is plumbing around two collectives on a communicator
// the caller hands over (or the library creates from an ncclUniqueId): the average of the parameter gradients and the
// sums behind ActNorm's data-dependent initialisation.  NCCL is loaded at run time (dlopen): a process that never
// touches these entry points needs no NCCL at all.
#pragma once
#include "common.cuh"
#include <vector>

namespace inb {

struct DpComm {
  void* comm = nullptr;       // ncclComm_t
  int nranks = 1, rank = 0;
  bool owned = false;         // created by inb_comm_create (destroyed with the handle)
  cudaStream_t st = nullptr;  // the stream the gradient all-reduces of an attached plan run on
  std::vector<cudaEvent_t> ev;
  int next_ev = 0;
  bool pending = false;       // collectives were enqueued on `st` during the current call
  long long calls = 0, bytes = 0;  // accounting: all-reduce calls issued / bytes reduced since creation
  cudaEvent_t event();
};

// sum over the ranks, in place, on stream st (count doubles)
void dp_allreduce_sum_f64(DpComm* dp, double* buf, size_t count, cudaStream_t st);
// One bucket: the tensors [first, first + n) of a pointer table averaged over the ranks.  Tensors that follow each
// other in memory (gap of `max_gap` elements at most - the caller's own alignment padding in the canonical flat layout,
// 0 otherwise) are reduced as one range.
void dp_allreduce_avg_bucket(DpComm* dp, float* const* ptrs, const long long* numel, int first, int n, long long max_gap,
                             cudaStream_t st);
// the comm stream waits for everything enqueued so far on `main` (and on `side`, when non-null)
void dp_fork(DpComm* dp, cudaStream_t main, cudaStream_t side);
// `main` waits for the collectives enqueued on the comm stream
void dp_join(DpComm* dp, cudaStream_t main);

}  // namespace inb
