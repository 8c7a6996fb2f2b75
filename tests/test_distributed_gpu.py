"""N > 1 on real GPUs (needs >= 2 visible devices; skipped otherwise): one process per GPU, the C-ABI collective
(cc_comm_create + in-place cc_allgather_prefix over NCCL) inside the serving loop, and the exactness claim of SURVEY 8d / 8e —
rank r's token ids at N = 2 equal the ids a single process computes for the same images, bit for bit, and every rank ends
up with the whole gathered prefix tensor."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from clipcap_b200.distributed import PrefixComm
        from clipcap_b200.pipeline import CaptionPipeline
        from oracle import synth
        from test_pipeline_gpu import _build
        encode_fn, model, vcfg, stop = _build(dev)
        B, EL = 4, 7
        comm = PrefixComm.from_process_group(dev)
        ok = comm.rank == rank and comm.world == world
        batches = [synth.pixels(world * B, vcfg.image_size, seed=90 + i) for i in range(3)]   # the job's global batches
        for sms in (0, 32):
            pipe = CaptionPipeline(encode_fn, model, B, vcfg.image_size, EL, stop, dev, comm=comm,
                                   prefix_dtype=torch.float16, partition_sms=sms)
            mine = [b[rank * B:(rank + 1) * B].to(dev) for b in batches]
            got = [(t.clone(), l.clone()) for t, l in pipe.run(mine, resident=True)]
            torch.cuda.synchronize()
            gathered = pipe.prefix_all.clone()   # after the last batch
            # single-process reference on this rank's device: the whole global batch, no collective
            solo = CaptionPipeline(encode_fn, model, world * B, vcfg.image_size, EL, stop, dev, prefix_dtype=torch.float16)
            want = [(t.clone(), l.clone()) for t, l in solo.run([b.to(dev) for b in batches], resident=True)]
            for (gt, gl), (wt, wl) in zip(got, want):
                ok &= torch.equal(gt, wt[rank * B:(rank + 1) * B]) and torch.equal(gl, wl[rank * B:(rank + 1) * B])
            full_prefix = model.transformer_mapper(encode_fn(batches[-1].to(dev)), out_dtype=torch.float16)
            ok &= torch.equal(gathered, full_prefix)   # every slot filled with exactly what a single process computes
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one_rank_exactly():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def test_single_rank_communicator(cuda_device):
    """nranks == 1: the communicator is created through NCCL like any other and the collective is the identity."""
    from clipcap_b200.distributed import PrefixComm
    comm = PrefixComm(0, 1, PrefixComm.make_unique_id(), cuda_device)
    x = torch.randn(3, 4, 8, device=cuda_device).half()
    y = x.clone()
    assert comm.all_gather_(y) is y and torch.equal(x, y) and comm.slot(y).data_ptr() == y.data_ptr()
    with pytest.raises(ValueError):
        comm.all_gather_(y.transpose(0, 1))
    comm.close()
