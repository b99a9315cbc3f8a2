"""Drop-in for `ladcast.pipelines.pipeline_AR.AutoRegressive2DPipeline` (reference pipelines/pipeline_AR.py:10-107):
the DPM-Solver++(2M) sampling loop — N denoiser calls with the scheduler update fused between them."""
from typing import List, Optional, Tuple, Union

import torch

from .. import _lib
from .utils import Fields2DPipelineOutput, randn_tensor


class AutoRegressive2DPipeline:
    model_cpu_offload_seq = "unet"

    def __init__(self, ar_model, scheduler, scheduler_step_kwargs: Optional[dict] = None):
        self.ar_model = ar_model
        self.scheduler = scheduler
        self.scheduler_step_kwargs = scheduler_step_kwargs or {}
        # the reference forwards these to scheduler.step (pipeline_AR.py:100); the fused step kernel is deterministic
        # and returns tensors, so only the no-op keys of EDMDPMSolverMultistepScheduler.step are accepted
        bad = [k for k in self.scheduler_step_kwargs if k not in ("generator", "return_dict")]
        if bad:
            raise NotImplementedError(f"scheduler_step_kwargs {bad} are not supported by the fused DPM-Solver++ step")

    @property
    def _execution_device(self):
        return self.ar_model.device

    @property
    def device(self):
        return self.ar_model.device

    def to(self, *a, **k):
        self.ar_model.to(*a, **k)
        return self

    def return_trajectory(self, *a, **k):
        raise NotImplementedError("This function is not implemented yet.")

    @torch.no_grad()
    def __call__(self, batch_size: int = 1, return_seq_len: int = 1, known_latents: torch.Tensor = None,
                 timestamps: Optional[torch.LongTensor] = None,
                 generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
                 num_inference_steps: int = 50, return_dict: bool = True,
                 do_edm_style: bool = True) -> Union[Fields2DPipelineOutput, Tuple]:
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(
                f"You have passed a list of generators of length {len(generator)}, but requested an effective batch"
                f" size of {batch_size}. Make sure the batch size matches the length of the generators.")
        assert known_latents is not None, "known_latents must be provided"
        if not do_edm_style:
            raise NotImplementedError("Only EDM style is supported for now")
        dev = self._execution_device
        shape = (batch_size, self.ar_model.config.out_channels, return_seq_len, *known_latents.shape[-2:])
        # x_T ~ N(0, 1): NOT multiplied by init_noise_sigma, exactly as the reference (pipeline_AR.py:77-82)
        image = randn_tensor(shape, generator=generator, device=dev, dtype=self.ar_model.dtype).contiguous()
        sch = self.scheduler
        sch.set_timesteps(num_inference_steps)
        n = len(sch.timesteps)
        lib = _lib.load()
        x0_prev = torch.empty_like(image)  # written by step 0 (first order: not read), read from step 1 on
        x_in = torch.empty_like(image)
        c_noise = sch.timesteps.to(dev, torch.float32)
        with self.ar_model.cached_conditioning(known_latents, timestamps, t_out=return_seq_len):
            c0 = sch.coefficients(0)
            _lib.check(lib.lc_sched_scale_input(_lib.ptr(image), _lib.ptr(x_in), image.numel(), c0["c_in"], _lib.stream()),
                       "lc_sched_scale_input")  # scale_model_input of step 0; later steps: fused into the step kernel
            for i in range(n):
                t = c_noise[i : i + 1]  # one c_noise for the whole batch (the reference expands it to (B,), :92)
                model_output = self.ar_model(x_in, t, known_latents, time_elapsed=timestamps, return_dict=False)[0]
                c = sch.coefficients(i)
                _lib.check(lib.lc_sched_dpmpp2m_step(
                    _lib.ptr(model_output), _lib.ptr(image), _lib.ptr(x0_prev), _lib.ptr(x_in) if i + 1 < n else None,
                    image.numel(), c["c_skip"], c["c_out"], c["a_x"], c["a_x0"], c["a_d"], c["c_in_next"],
                    _lib.stream()), "lc_sched_dpmpp2m_step")
                if sch.lower_order_nums < sch.config.solver_order:
                    sch.lower_order_nums += 1
        if not return_dict:
            return (image,)
        return Fields2DPipelineOutput(fields=image)
