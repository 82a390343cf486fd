"""2/4/8-GPU check of the head-parallel single-stream path (run under torch.distributed.run, or spawned one
process per rank by bench.py): the sharded pipeline must reproduce the single-GPU CUDA pipeline on the same
weights / inputs and is timed against it.  Rank 0 writes a JSON summary (--out, default
gpurun_out/ulysses_P<P>_graph<g>[_fp8].json).

  --timeline 1     one extra eager pass with CUDA events around the five phases of every block (dense work, the
                   two producing kernels with their NVLink stores, the two peer barriers), per rank, plus the
                   algorithmic NVLink bytes per layer and - when NVML exposes them - the measured link counters
  --switch-race 1  interactive pipeline with one prompt switch; the last rank sleeps on the host right before the
                   KV-recache, so a faster peer reaches layer 0 of the recache forward (which stores K / V into the
                   peers' rings) while the slow rank is still zeroing its ring: the start-of-forward barrier must
                   make the result equal to the single-GPU interactive pipeline
  --fp8 1          W8A8 (e4m3) block linears on both sides
"""
import argparse
import json
import os
import sys
import time
import types

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longlive_b200 import synth  # noqa: E402
from longlive_b200.model import CausalWanModel  # noqa: E402
from longlive_b200.pipeline import CausalInferencePipeline, InteractiveCausalInferencePipeline  # noqa: E402
from longlive_b200.ulysses import UlyssesCausalWanModel  # noqa: E402
from longlive_b200.wrapper import WanDiffusionWrapper  # noqa: E402


def nvlink_kib(index: int):
    """(tx KiB, rx KiB) summed over the links of GPU `index`, or None when NVML does not expose the counters."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX,
                                                   pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                raise RuntimeError("field not available")
            out.append(int(v.value.ullVal))
        return tuple(out)
    except Exception:
        pass
    try:  # fallback: nvidia-smi's per-link data counters
        import re
        import subprocess
        txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True,
                             timeout=20).stdout
        tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", txt))
        rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", txt))
        return (tx, rx) if (tx or rx) else None
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=15)
    ap.add_argument("--layers", type=int, default=30)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--fp8", type=int, default=0)
    ap.add_argument("--timeline", type=int, default=0)
    ap.add_argument("--switch-race", type=int, default=0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    class MK(dict):
        __getattr__ = dict.get
    args = types.SimpleNamespace(denoising_step_list=[1000, 750, 500, 250], warp_denoising_step=True,
                                 num_frame_per_block=3, context_noise=0, global_sink=False,
                                 model_kwargs=MK(local_attn_size=12, sink_size=3, timestep_shift=5.0))
    prompts = [synth.prompt_embeds(100 + i).to(dev) for i in range(2)]
    noise = synth.latent_noise(0, a.frames).to(dev)

    def make_pipe(model, interactive=False, slow_rank=None):
        gen = WanDiffusionWrapper(model=model, timestep_shift=5.0)
        cls = InteractiveCausalInferencePipeline if interactive else CausalInferencePipeline
        pipe = cls(args, dev, generator=gen,
                   text_encoder=lambda text_prompts: {"prompt_embeds": prompts[int(text_prompts[0])]})
        if slow_rank is not None and rank == slow_rank:
            orig = pipe._recache_after_switch

            def delayed(*x, **k):
                torch.cuda.synchronize()
                time.sleep(0.3)  # the peers are already inside the recache forward when this rank zeroes its ring
                return orig(*x, **k)
            pipe._recache_after_switch = delayed
        return pipe

    def run(pipe, interactive=False):
        outs, times = None, []
        for it in range(2):  # second pass is timed (first includes lazy init / graph capture)
            torch.manual_seed(1234)  # identical re-noise draws on every rank and every pass
            torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
            if interactive:
                _, lat = pipe.inference(noise, text_prompts_list=[["0"], ["1"]], switch_frame_indices=[a.frames // 2],
                                        return_latents=True)
            else:
                _, lat = pipe.inference(noise, ["0"], return_latents=True)
            torch.cuda.synchronize(); dist.barrier(); times.append(time.time() - t0)
            outs = lat
        return outs, times[-1]

    inter = bool(a.switch_race)
    base = CausalWanModel(local_attn_size=12, sink_size=3, num_layers=a.layers)
    synth.random_init_(base, 0)
    base.fp8_linears = bool(a.fp8)
    base = base.to(dev).to(torch.bfloat16)
    base.use_cuda_graph = bool(a.graph)
    ref_lat, t_ref = run(make_pipe(base, inter), inter)
    sd = base.state_dict()
    del base
    torch.cuda.empty_cache()

    par = UlyssesCausalWanModel(local_attn_size=12, sink_size=3, num_layers=a.layers)
    par.load_state_dict(sd)
    par.fp8_linears = bool(a.fp8)
    par = par.to(dev).to(torch.bfloat16)
    par.setup_parallel(use_cuda_graph=bool(a.graph))
    pipe = make_pipe(par, inter, slow_rank=world - 1 if inter else None)
    lat, t_par = run(pipe, inter)

    errs = [(((lat[:, c:c + 3].float() - ref_lat[:, c:c + 3].float()).norm() /
              ref_lat[:, c:c + 3].float().norm()).item()) for c in range(0, a.frames, 3)]
    res = {"P": world, "frames": a.frames, "layers": a.layers, "fp8_linears": bool(a.fp8),
           "head_map": "round-robin (uneven)" if par.round_robin else "contiguous",
           "heads_per_rank": [len(range(r, par.num_heads, world)) if par.round_robin else par.hp for r in range(world)],
           "switch_race": inter, "rel_l2_per_chunk_vs_single_gpu": errs,
           "cuda_graph": bool(a.graph), "seconds_single_gpu": t_ref, "seconds_head_parallel": t_par,
           "speedup": t_ref / t_par, "efficiency": t_ref / t_par / world,
           "fps_single": 4 * a.frames / t_ref, "fps_parallel": 4 * a.frames / t_par}

    if a.timeline:
        # one eager pass with events; graphs dropped first (they would bypass the marks)
        par._graphs.clear()
        par.use_cuda_graph = False
        torch.manual_seed(1234)
        pipe.inference(noise[:, :6], ["0"], return_latents=True)        # eager warm-up
        link0 = nvlink_kib(local)
        par.start_timeline()
        n_fwd0 = par.kernel_launches
        torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
        pipe.inference(noise, ["0"], return_latents=True)
        torch.cuda.synchronize(); t_eager = time.time() - t0
        phases = par.phase_ms()
        link1 = nvlink_kib(local)
        forwards = 5 * (a.frames // 3)
        mine = {"rank": rank, "phase_ms_per_forward": {k: v / forwards for k, v in phases.items()},
                "eager_seconds": t_eager,
                "nvlink_tx_rx_MiB_per_forward": ([(b - a_) / 1024 / forwards for a_, b in zip(link0, link1)]
                                                 if link0 and link1 else None)}
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        L, C_ = 3 * 1560, 1536
        res["timeline"] = {
            "per_rank": allr,
            "algorithmic_nvlink_MiB_per_forward_per_rank": a.layers * (4 * (L // world) * C_ * 2) * (world - 1) / world / 2 ** 20,
            "note": "per rank and layer a token shard sends q, k, v (3 x L/P x 1536 x 2 B) and receives its attention rows "
                    "(L/P x 1536 x 2 B) of which the (P-1)/P share crosses NVLink; steady-state chunks (L = 4680)"}
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        out = a.out or (f"gpurun_out/ulysses_P{world}_graph{a.graph}" + ("_fp8" if a.fp8 else "") +
                        ("_switch_race" if inter else "") + ".json")
        json.dump(res, open(out, "w"), indent=1)
        print(json.dumps(res))
    # bf16: the two paths differ by summation order only (measured 6.4e-3, the chaos floor of the random-init stack).
    # W8A8: the same reordering noise is amplified by the e4m3 re-quantisation of every activation (measured 2.1e-2; the
    # fp8 option itself sits 3.1e-2 from the bf16 path, tests/test_fp8_gpu.py), hence the wider gate.
    ok = max(errs) < (5e-2 if a.fp8 else 1e-2)
    # captured graphs hold NCCL work: drop them before tearing the communicator down, and leave
    # through os._exit so a stuck communicator destructor can never hang the box
    par._graphs.clear()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
