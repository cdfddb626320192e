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


def benchmark(model: torch.nn.Module, device=0, input_size=(3, 16, 224, 224), batch_size: int = 8,
              runs: int = 40, throw_out: float = 0.25, use_bf16: bool = False, verbose: bool = False) -> float:
    """Throughput in clips/s (tome/utils.py:15-80: first ``throw_out`` of the runs discarded),
    timed on the device with CUDA events."""
    dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device
    model = model.eval().to(dev)
    x = torch.rand(batch_size, *input_size, device=dev)
    warm = int(runs * throw_out)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_bf16):
        for i in range(runs):
            if i == warm:
                torch.cuda.synchronize(dev)
                start.record()
            model([x])
        end.record()
    torch.cuda.synchronize(dev)
    thr = (runs - warm) * batch_size / (start.elapsed_time(end) * 1e-3)
    if verbose:
        print(f"Throughput: {thr:.2f} clips/s")
    return thr
