"""Call-compatible stand-in for the reference's ``src/simulator/config.py::jax_init`` (config.py:73-153).

The reference configures XLA there (fake host devices, x64 switch, allocator) and must do so before ``import jax``.
Nothing of that exists on this path -- arithmetic is FP64 unless ``precision='fp32'`` is asked for, one process
drives one GPU -- so a script that starts with ``config.jax_init(...)`` keeps working and is told what it runs on."""


def jax_init(force_device=None, core_limit=None, extra_info=False, disable_python_multithreading=True, enable_x64=False,
             debugging=False):
    import torch
    if force_device == "cpu":
        raise RuntimeError("synthpy_b200 has no CPU path (force_device='cpu')")
    if not torch.cuda.is_available():
        raise RuntimeError("synthpy_b200 needs a CUDA device: the hot path has no CPU fallback")
    dev = torch.cuda.current_device()
    p = torch.cuda.get_device_properties(dev)
    print(f"synthpy_b200: cuda:{dev} {p.name}, {p.multi_processor_count} SMs, {p.total_memory / 2 ** 30:.0f} GiB; "
          f"ray state and arithmetic are float64 (enable_x64 is implied; precision='fp32' selects the float32 mode)")
    if extra_info:
        print(f"  torch {torch.__version__}, CUDA runtime {torch.version.cuda}, {torch.cuda.device_count()} visible device(s)")
    return [torch.device("cuda", i) for i in range(torch.cuda.device_count())]
