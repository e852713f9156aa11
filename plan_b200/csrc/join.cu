// join.cu -- hash join pipelines (placeholder until the bucketized table lands).
#include "pipeline.hpp"

namespace pg {

int build_join_agg(pg_plan *, const Node &, const Node &, std::unique_ptr<Pipeline> *)
{
    PG_FAIL(PG_EUNSUPPORTED, "join pipelines are not built yet");
}

}  // namespace pg
