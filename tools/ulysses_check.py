"""2/4-GPU check of the head-parallel single-stream path (run under torch.distributed.run):
the sharded pipeline must reproduce the single-GPU CUDA pipeline on the same weights / inputs, and is
timed against it.  Writes gpurun_out/ulysses_P<P>.json on rank 0."""
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
from longlive_b200.pipeline import CausalInferencePipeline  # noqa: E402
from longlive_b200.ulysses import UlyssesCausalWanModel  # noqa: E402
from longlive_b200.wrapper import WanDiffusionWrapper  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=15)
    ap.add_argument("--layers", type=int, default=30)
    ap.add_argument("--graph", type=int, default=0)
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
    prompt = synth.prompt_embeds(100).to(dev)
    noise = synth.latent_noise(0, a.frames).to(dev)

    def run(model, tag):
        gen = WanDiffusionWrapper(model=model, timestep_shift=5.0)
        pipe = CausalInferencePipeline(args, dev, generator=gen,
                                       text_encoder=lambda text_prompts: {"prompt_embeds": prompt})
        outs, times = None, []
        for it in range(2):  # second pass is timed (first includes lazy init)
            torch.manual_seed(1234)  # identical re-noise draws on every rank and every pass
            torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
            _, lat = pipe.inference(noise, ["p"], return_latents=True)
            torch.cuda.synchronize(); dist.barrier(); times.append(time.time() - t0)
            outs = lat
        return outs, times[-1]

    base = CausalWanModel(local_attn_size=12, sink_size=3, num_layers=a.layers)
    synth.random_init_(base, 0)
    base = base.to(dev).to(torch.bfloat16)
    base.use_cuda_graph = bool(a.graph)
    ref_lat, t_ref = run(base, "single")
    sd = base.state_dict()
    del base
    torch.cuda.empty_cache()

    par = UlyssesCausalWanModel(local_attn_size=12, sink_size=3, num_layers=a.layers)
    par.load_state_dict(sd)
    par = par.to(dev).to(torch.bfloat16)
    par.setup_parallel(use_cuda_graph=bool(a.graph))
    lat, t_par = run(par, "ulysses")

    errs = [(((lat[:, c:c + 3].float() - ref_lat[:, c:c + 3].float()).norm() /
              ref_lat[:, c:c + 3].float().norm()).item()) for c in range(0, a.frames, 3)]
    res = {"P": world, "frames": a.frames, "layers": a.layers, "rel_l2_per_chunk_vs_single_gpu": errs,
           "cuda_graph": bool(a.graph), "seconds_single_gpu": t_ref, "seconds_head_parallel": t_par,
           "speedup": t_ref / t_par, "fps_single": 4 * a.frames / t_ref, "fps_parallel": 4 * a.frames / t_par}
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open(f"gpurun_out/ulysses_P{world}_graph{a.graph}.json", "w"), indent=1)
        print(json.dumps(res))
    ok = max(errs) < 1e-2
    # captured graphs hold NCCL work: drop them before tearing the communicator down, and leave
    # through os._exit so a stuck communicator destructor can never hang the box
    par._graphs.clear()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
