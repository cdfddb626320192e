"""``tome.utils`` of the reference (tome/utils.py): the r-schedule parser the patches use,
and the throughput helper's warm-up / throw-out protocol restated for CUDA events."""
import time
from typing import List, Tuple, Union

import torch


def parse_r(num_layers: int, r: Union[List[int], Tuple[int, float], int]) -> List[int]:
    """Constant r, (r, inflection) or per-layer list -> per-layer list (tome/utils.py:83-108).

    inflection +1 trends upward, -1 downward ("decreasing schedule"), 0 is constant."""
    inflect = 0
    if isinstance(r, list):
        if len(r) < num_layers:
            r = r + [0] * (num_layers - len(r))
        return list(r)
    elif isinstance(r, tuple):
        r, inflect = r
    min_val = int(r * (1.0 - inflect))
    max_val = 2 * r - min_val
    step = (max_val - min_val) / (num_layers - 1)
    return [int(min_val + step * i) for i in range(num_layers)]


def benchmark(model: torch.nn.Module, device=0, input_size: Tuple[int] = (3, 224, 224), batch_size: int = 64,
              runs: int = 40, throw_out: float = 0.25, use_fp16: bool = False, verbose: bool = False, *,
              use_bf16: bool = False, as_pathways: bool = False) -> float:
    """Throughput helper with the reference's signature and protocol (tome/utils.py:15-80): random inputs of
    ``(batch_size, *input_size)``, the first ``throw_out`` of ``runs`` discarded, images/s for a 3-d
    ``input_size`` and frames/s (batch x input_size[1]) for a 4-d one, host clock around a device synchronise.
    Additions (keyword-only): ``use_bf16`` autocasts to bfloat16 instead of float16; ``as_pathways`` calls
    ``model([x])``, the list-of-pathways input the SlowFast-style video wrappers take."""
    if not isinstance(device, torch.device):
        device = torch.device(device)
    is_cuda = device.type == "cuda"
    model = model.eval().to(device)
    x = torch.rand(batch_size, *input_size, device=device)
    if use_fp16:
        x = x.half()
    warm_up = int(runs * throw_out)
    per_run = batch_size * (input_size[1] if len(input_size) == 4 else 1)
    lowp = torch.bfloat16 if use_bf16 else torch.float16
    total, start = 0, time.time()
    with torch.autocast(device.type, dtype=lowp, enabled=use_fp16 or use_bf16), torch.no_grad():
        for i in range(runs):
            if i == warm_up:
                if is_cuda:
                    torch.cuda.synchronize(device)
                total, start = 0, time.time()
            model([x] if as_pathways else x)
            total += per_run
    if is_cuda:
        torch.cuda.synchronize(device)
    throughput = total / (time.time() - start)
    if verbose:
        print(f"Throughput: {throughput:.2f} im/s")
    return throughput
