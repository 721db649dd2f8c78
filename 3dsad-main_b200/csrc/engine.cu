// Host side of the pipelined executor (engine.PipelinedHotPath, SURVEY.md section 8(e) / VERDICT r1 item 8): one batch
// -- optional event wait, input copies, the slot's CUDA graph, result copies, completion event -- queued by ONE C call.
// Through PyTorch the same sequence is ~10 Python-level calls (stream context, three copy_, replay, two copy_, event
// record) and costs more host time per batch than the GPU needs to start it, which staggers the fill of a pipeline
// with 32 batches in flight.  (No reference file exists to cite: /root/reference is README.md:1-2 only.)
#include "sad_common.cuh"

extern "C" int sad_engine_submit(const sad_submit_desc* d, sad_stream_t stream) {
  SAD_REQUIRE(d && d->graph_exec, "engine_submit: null descriptor / graph");
  SAD_REQUIRE(d->n_in >= 0 && d->n_in <= SAD_SUBMIT_MAX_COPIES && d->n_out >= 0 && d->n_out <= SAD_SUBMIT_MAX_COPIES,
              "engine_submit: bad copy counts %d / %d", d->n_in, d->n_out);
  cudaStream_t st = (cudaStream_t)stream;
  if (d->wait_event) SAD_CUDA_OK(cudaStreamWaitEvent(st, (cudaEvent_t)d->wait_event, 0));
  for (int i = 0; i < d->n_in; ++i) {
    SAD_REQUIRE(d->in_dst[i] && d->in_src[i], "engine_submit: null input copy %d", i);
    SAD_CUDA_OK(cudaMemcpyAsync(d->in_dst[i], d->in_src[i], d->in_bytes[i], cudaMemcpyDefault, st));
  }
  SAD_CUDA_OK(cudaGraphLaunch((cudaGraphExec_t)d->graph_exec, st));
  for (int i = 0; i < d->n_out; ++i) {
    SAD_REQUIRE(d->out_dst[i] && d->out_src[i], "engine_submit: null output copy %d", i);
    SAD_CUDA_OK(cudaMemcpyAsync(d->out_dst[i], d->out_src[i], d->out_bytes[i], cudaMemcpyDefault, st));
  }
  if (d->done_event) SAD_CUDA_OK(cudaEventRecord((cudaEvent_t)d->done_event, st));
  return SAD_OK;
}
