"""N > 1 host path on CPU: two gloo ranks, each with its own stream; only a barrier and a MAX
reduction cross ranks (no data-path collective).  Also checks that two ranks driving the pipeline
host logic produce independent, correctly-indexed call sequences."""
import os
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from longlive_b200 import multistream


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cpu")
        multistream.barrier(dev)
        streams = multistream.stream_assignment(5, rank, world)
        # rank 1 is "slower": the job time is the max over ranks
        agg = multistream.aggregate_throughput(units_per_rank=84.0, elapsed_ms_local=1000.0 * (rank + 1), device=dev)
        # drive the real pipeline host logic with a recording generator on this rank's stream
        from tests.test_pipeline_host_cpu import FakeGenerator, _args
        from longlive_b200.pipeline import CausalInferencePipeline
        gen = FakeGenerator()
        pipe = CausalInferencePipeline(_args(), dev, generator=gen,
                                       text_encoder=lambda text_prompts: {"id": text_prompts[0]})
        noise_seed, prompt_seed = multistream.stream_seeds(streams[0])
        g = torch.Generator().manual_seed(noise_seed)
        noise = torch.randn(1, 6, 16, 4, 4, generator=g).to(torch.bfloat16)
        pipe.inference(noise, [f"prompt-{prompt_seed}"])
        multistream.barrier(dev)
        out[rank] = {"streams": streams, "agg": agg, "calls": len(gen.calls),
                     "prompt": gen.calls[0]["prompt"], "noise0": float(noise.flatten()[0])}
    finally:
        dist.destroy_process_group()


def test_two_ranks_independent_streams():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["streams"] == [0, 2, 4] and r1["streams"] == [1, 3]
    # value = all ranks' units / slowest rank's time
    for r in (r0, r1):
        assert r["agg"]["world"] == 2 and r["agg"]["ms"] == pytest.approx(2000.0)
        assert r["agg"]["value"] == pytest.approx(2 * 84.0 / 2.0)
    assert r0["calls"] == r1["calls"] == 10
    assert (r0["prompt"], r1["prompt"]) == ("prompt-100", "prompt-101")
    assert r0["noise0"] != r1["noise0"]


def test_single_process_fallbacks():
    assert multistream.stream_assignment(3, 0, 1) == [0, 1, 2]
    agg = multistream.aggregate_throughput(84.0, 500.0, torch.device("cpu"))
    assert agg == {"world": 1, "ms": 500.0, "value": 168.0}
