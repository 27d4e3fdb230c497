// pipeline_poly.cu -- rational-ratio (polyphase) fast path of the fused pipeline.
#include "pipeline.cuh"

int aukit_pipeline_poly_try(aukit_ctx *, const pipe_args &, const aukit_pipeline_desc *, bool) { return 0; }
