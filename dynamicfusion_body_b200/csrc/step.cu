// step.cu -- one frame of the warped projective TSDF path (a3) as ONE CUDA-graph launch.
//
// The reference's frame loop (test.py:116-137) calls `updateTSDF` once per frame; on a z-slab of a volume sharded over
// 8 GPUs that call is ~0.1 ms of device work (SURVEY 7 "multi-GPU scaling is latency-bound"), so the per-frame host work
// has to shrink to a single launch.  A dfb_frame_step captures
//     [root: upload of the frame's node transforms] -> [NCCL broadcast of the transforms] -> dfb_nodes_pack
//     -> dfb_tsdf_update_projective(DFB_MODE_HYBRID)  -> [read-back of the 8 frame counters to pinned host memory]
// and, as a concurrent branch of the same graph, the upload + broadcast of the NEXT frame's sensor data into the other
// half of a double buffer (depth does not depend on the fusion result; the transforms do).  The first call captures the
// launches from the stream; a call with byte-identical arguments replays the executable graph (one cudaGraphLaunch); a
// call whose arguments changed (a new global rigid dq, other buffers) re-captures and updates the executable graph in
// place (cudaGraphExecUpdate), re-instantiating only when the topology changed.
#include <string.h>

#include "common.h"

struct dfb_frame_step {
    cudaGraphExec_t exec = nullptr;
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool have_key = false;
    dfb_volume vol;
    dfb_warpfield wf;
    dfb_views views;
    dfb_workspace ws;
    dfb_frame_io io;
    double tdist, wmax;
    int64_t captures = 0, updates = 0, replays = 0, direct = 0;
    size_t graph_nodes = 0;
};

namespace {

int issue(dfb_frame_step* st, const dfb_volume* vol, const dfb_warpfield* wf, const dfb_views* views, double tdist, double wmax,
          const dfb_workspace* ws, const dfb_frame_io* io, cudaStream_t s, bool forked) {
    const bool is_root = !io || !io->comm || dfb_comm_rank(io->comm) == io->root;
    const bool pre_root = !io || !io->comm_prefetch || dfb_comm_rank(io->comm_prefetch) == io->root;
    const bool prefetch = io && io->prefetch_dst && io->prefetch_bytes > 0 && (io->prefetch_src || io->comm_prefetch);
    if (prefetch) {
        cudaStream_t b = forked ? st->side : s;
        if (forked) {
            DFB_CUDA(cudaEventRecord(st->fork, s));
            DFB_CUDA(cudaStreamWaitEvent(st->side, st->fork, 0));
        }
        if (io->prefetch_src && pre_root)
            DFB_CUDA(cudaMemcpyAsync(io->prefetch_dst, io->prefetch_src, (size_t)io->prefetch_bytes, cudaMemcpyDefault, b));
        if (io->comm_prefetch)
            if (int r = dfb_comm_broadcast(io->comm_prefetch, io->prefetch_dst, io->prefetch_bytes, io->root, b)) return r;
        if (forked) DFB_CUDA(cudaEventRecord(st->join, st->side));
    }
    float* node_dq = const_cast<float*>(wf->node_dq);
    if (io && io->dq_src && is_root)
        DFB_CUDA(cudaMemcpyAsync(node_dq, io->dq_src, (size_t)wf->n_nodes * 8 * sizeof(float), cudaMemcpyDefault, s));
    if (io && io->comm)
        if (int r = dfb_comm_broadcast(io->comm, node_dq, (int64_t)wf->n_nodes * 8 * sizeof(float), io->root, s)) return r;
    if (int r = dfb_nodes_pack(wf->node_pos, wf->node_dq, wf->node_w, wf->n_nodes, const_cast<float*>(wf->node_rec), s)) return r;
    if (int r = dfb_tsdf_update_projective(vol, wf, views, tdist, wmax, DFB_MODE_HYBRID, ws, nullptr, nullptr, s)) return r;
    if (io && io->counters_host)
        DFB_CUDA(cudaMemcpyAsync(io->counters_host, ws->counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    if (prefetch && forked) DFB_CUDA(cudaStreamWaitEvent(s, st->join, 0));
    return DFB_OK;
}

}  // namespace

extern "C" int dfb_frame_step_create(dfb_frame_step** out) {
    DFB_REQUIRE(out, "null pointer");
    dfb_frame_step* st = new dfb_frame_step;
    cudaError_t e = cudaStreamCreateWithFlags(&st->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&st->fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&st->join, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        dfb_frame_step_destroy(st);
        return dfb::check_cuda(e, "dfb_frame_step_create");
    }
    *out = st;
    return DFB_OK;
}

extern "C" void dfb_frame_step_destroy(dfb_frame_step* st) {
    if (!st) return;
    if (st->exec) cudaGraphExecDestroy(st->exec);
    if (st->fork) cudaEventDestroy(st->fork);
    if (st->join) cudaEventDestroy(st->join);
    if (st->side) cudaStreamDestroy(st->side);
    delete st;
}

extern "C" int dfb_frame_step_run(dfb_frame_step* st, const dfb_volume* vol, const dfb_warpfield* wf, const dfb_views* views, double tdist,
                                  double wmax, const dfb_workspace* ws, const dfb_frame_io* io, dfb_stream_t stream) {
    DFB_REQUIRE(st && vol && wf && views && ws, "null pointer");
    DFB_REQUIRE(wf->node_rec && wf->node_pos && wf->node_dq && wf->node_w && wf->n_nodes > 0, "node arrays are null");
    cudaStream_t s = (cudaStream_t)stream;
    dfb_frame_io io0;
    memset(&io0, 0, sizeof(io0));
    if (io) io0 = *io;
    // the legacy default stream cannot be captured: plain launches
    if (s == nullptr || s == cudaStreamLegacy) {
        ++st->direct;
        return issue(st, vol, wf, views, tdist, wmax, ws, &io0, s, false);
    }
    const bool same = st->have_key && st->exec && memcmp(&st->vol, vol, sizeof(*vol)) == 0 && memcmp(&st->wf, wf, sizeof(*wf)) == 0 &&
                      memcmp(&st->views, views, sizeof(*views)) == 0 && memcmp(&st->ws, ws, sizeof(*ws)) == 0 &&
                      memcmp(&st->io, &io0, sizeof(io0)) == 0 && st->tdist == tdist && st->wmax == wmax;
    if (!same) {
        DFB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
        const int r = issue(st, vol, wf, views, tdist, wmax, ws, &io0, s, true);
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(s, &g);
        if (r != DFB_OK) {
            if (g) cudaGraphDestroy(g);
            return r;
        }
        DFB_CUDA(e);
        ++st->captures;
        bool need_instantiate = st->exec == nullptr;
        if (st->exec) {
            cudaGraphExecUpdateResultInfo info;
            if (cudaGraphExecUpdate(st->exec, g, &info) == cudaSuccess) {
                ++st->updates;
            } else {
                (void)cudaGetLastError();
                cudaGraphExecDestroy(st->exec);
                st->exec = nullptr;
                need_instantiate = true;
            }
        }
        if (need_instantiate) {
            const cudaError_t ei = cudaGraphInstantiate(&st->exec, g, 0);
            if (ei != cudaSuccess) {
                cudaGraphDestroy(g);
                st->exec = nullptr;
                st->have_key = false;
                return dfb::check_cuda(ei, "cudaGraphInstantiate");
            }
        }
        cudaGraphGetNodes(g, nullptr, &st->graph_nodes);
        cudaGraphDestroy(g);
        st->vol = *vol; st->wf = *wf; st->views = *views; st->ws = *ws; st->io = io0; st->tdist = tdist; st->wmax = wmax;
        st->have_key = true;
    } else {
        ++st->replays;
    }
    DFB_CUDA(cudaGraphLaunch(st->exec, s));
    return DFB_OK;
}

extern "C" int dfb_frame_step_stats(const dfb_frame_step* st, int64_t out[5]) {
    DFB_REQUIRE(st && out, "null pointer");
    out[0] = st->captures; out[1] = st->updates; out[2] = st->replays; out[3] = st->direct; out[4] = (int64_t)st->graph_nodes;
    return DFB_OK;
}
