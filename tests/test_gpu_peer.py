"""NVLink peer exchange (pft_tracker_peer_*): R processes, one tracker rank each, exchange the crop box and
the raw weights of weight() through CUDA-IPC windows written by the producing kernels themselves.  The
sharded run must equal the unsharded one bit for bit, over several compute() calls (CUDA-graph replays
included).  With fewer GPUs than ranks the ranks share GPU 0 (the driver time-slices the processes), so the
test also runs on a one-GPU box."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_PART, N_MAX, FRAMES = 101, 160, 4


def _tracker(kld, shard=None, device=0):
    from pcl_tracking_b200 import pcl
    from tests import util
    ctx = pcl.Context(device)
    scene, model, centre = util.small_case(77, n_scene=4000, n_model=260)
    g = (pcl.KLDAdaptiveParticleFilterOMPTracker if kld else pcl.ParticleFilterOMPTracker)(16, ctx=ctx)
    pcl.configure_like_reference(g, coherence_cls=pcl.NearestPairPointCloudCoherence, particle_num=N_PART, max_particle_num=N_MAX,
                                 use_hsv=True, iteration_num=2)
    if kld:
        g.setEpsilon(0.2)
        g.setBinSize([0.1] * 6)
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = centre
    g.setTrans(m)
    g.seed(99)
    if shard:
        g.setShard(*shard)
    g.setReferenceCloud(model)
    cloud = pcl.PointCloud(scene, ctx=ctx)
    g.setInputCloud(cloud)
    return ctx, g, cloud


def _run(g):
    out = []
    for _ in range(FRAMES):
        g.compute()
        out.append((g.getParticles().copy(), g.rawWeights().copy(), np.array(g.getResult().tolist(), dtype=np.float32)))
    return out


def _rank_main(rank, nranks, kld, conns, result_q):
    try:
        sys.path.insert(0, ROOT)
        import torch  # noqa: F401  (device count)
        device = rank if torch.cuda.device_count() >= nranks else 0
        ctx, g, cloud = _tracker(kld, (nranks, rank), device)
        mine = g.peerExport()
        # all-gather of the handles through the parent
        conns[rank].send(mine)
        handles = conns[rank].recv()
        g.peerAttach(handles)
        conns[rank].send("attached")
        assert conns[rank].recv() == "go"
        out = _run(g)
        replays = g.graphReplays()
        g.peerDetach()
        result_q.put((rank, out, replays, None))
    except Exception as e:  # pragma: no cover
        import traceback
        result_q.put((rank, None, 0, traceback.format_exc() + repr(e)))


@pytest.mark.parametrize("nranks,kld", [(2, False), (3, True)])
def test_peer_exchange_equals_unsharded(nranks, kld):
    mpc = mp.get_context("spawn")
    pipes = [mpc.Pipe() for _ in range(nranks)]
    q = mpc.Queue()
    procs = [mpc.Process(target=_rank_main, args=(r, nranks, kld, [p[1] for p in pipes], q)) for r in range(nranks)]
    for p in procs:
        p.start()
    try:
        handles = []
        for r in range(nranks):
            assert pipes[r][0].poll(180), "rank %d did not export its window" % r
            handles.append(pipes[r][0].recv())
        for r in range(nranks):
            pipes[r][0].send(handles)
        for r in range(nranks):
            assert pipes[r][0].poll(120) and pipes[r][0].recv() == "attached"
        for r in range(nranks):
            pipes[r][0].send("go")
        results = {}
        for _ in range(nranks):
            rank, out, replays, err = q.get(timeout=300)
            assert err is None, "rank %d failed:\n%s" % (rank, err)
            results[rank] = (out, replays)
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    # the unsharded run, in this process
    ctx, g, cloud = _tracker(kld)
    want = _run(g)
    for rank in range(nranks):
        out, replays = results[rank]
        assert replays >= 1, "steady-state frames must replay the CUDA graph with the peer exchange captured in it"
        for f in range(FRAMES):
            assert np.array_equal(out[f][0].view(np.uint32), want[f][0].view(np.uint32)), "particles differ: rank %d frame %d" % (rank, f)
            assert np.array_equal(out[f][1].view(np.uint32), want[f][1].view(np.uint32)), "raw weights differ: rank %d frame %d" % (rank, f)
            assert np.array_equal(out[f][2].view(np.uint32), want[f][2].view(np.uint32)), "result differs: rank %d frame %d" % (rank, f)
