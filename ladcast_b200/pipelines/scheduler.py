"""Karras/EDM DPM-Solver++ scheduler — duck type of `diffusers.EDMDPMSolverMultistepScheduler` as the reference
uses it (evaluate/pred_rollout.py:49-52, pipelines/pipeline_AR.py:85-102, pipelines/edm_sampler.py:56-110): same
attribute / method surface (`config`, `set_timesteps`, `timesteps`, `sigmas`, `scale_model_input`, `step`,
`precondition_inputs/noise/outputs`, `init_noise_sigma`).  The tensor update of `step` runs in the fused CUDA kernel
`lc_sched_dpmpp2m_step`; the handful of scalar coefficients are computed on the host in fp32, in the same order of
operations as diffusers, so they are bit-identical to the reference's 0-dim tensor arithmetic."""
import torch

from .. import _lib
from ..models.modeling import Config


class EDMDPMSolverMultistepScheduler:
    order = 1

    def __init__(self, sigma_min: float = 0.002, sigma_max: float = 80.0, sigma_data: float = 0.5,
                 sigma_schedule: str = "karras", num_train_timesteps: int = 1000, prediction_type: str = "epsilon",
                 rho: float = 7.0, solver_order: int = 2, thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, sample_max_value: float = 1.0,
                 algorithm_type: str = "dpmsolver++", solver_type: str = "midpoint", lower_order_final: bool = True,
                 euler_at_final: bool = False, final_sigmas_type: str = "zero"):
        kw = dict(locals())
        kw.pop("self")
        self.config = Config(kw)
        if (algorithm_type, solver_type, sigma_schedule) != ("dpmsolver++", "midpoint", "karras") or thresholding:
            raise NotImplementedError("only the reference's default EDM DPM-Solver++ configuration is implemented")
        if solver_order not in (1, 2) or prediction_type != "epsilon":
            raise NotImplementedError("solver_order must be 1 or 2 and prediction_type 'epsilon'")
        self.num_inference_steps = None
        self.set_timesteps(num_train_timesteps)
        self.num_inference_steps = None

    # -- schedule
    def _karras(self, ramp):
        lo, hi = self.config.sigma_min ** (1 / self.config.rho), self.config.sigma_max ** (1 / self.config.rho)
        return (hi + ramp * (lo - hi)) ** self.config.rho

    def set_timesteps(self, num_inference_steps=None, device=None):
        self.num_inference_steps = num_inference_steps
        sigmas = self._karras(torch.linspace(0, 1, num_inference_steps)).to(torch.float32)
        self.timesteps = self.precondition_noise(sigmas).to(device) if device is not None else self.precondition_noise(sigmas)
        last = self.config.sigma_min if self.config.final_sigmas_type == "sigma_min" else 0.0
        self.sigmas = torch.cat([sigmas, torch.tensor([last], dtype=torch.float32)])  # kept on CPU
        self._x0_prev = None
        self.lower_order_nums = 0
        self._step_index = None

    @property
    def init_noise_sigma(self):
        return (self.config.sigma_max**2 + 1) ** 0.5

    @property
    def step_index(self):
        return self._step_index

    # -- EDM preconditioning
    def precondition_inputs(self, sample, sigma):
        return sample * (1 / ((sigma**2 + self.config.sigma_data**2) ** 0.5))

    def precondition_noise(self, sigma):
        if not isinstance(sigma, torch.Tensor):
            sigma = torch.tensor([sigma])
        return 0.25 * torch.log(sigma)

    def precondition_outputs(self, sample, model_output, sigma):
        sd = self.config.sigma_data
        c_skip = sd**2 / (sigma**2 + sd**2)
        c_out = sigma * sd / (sigma**2 + sd**2) ** 0.5
        return c_skip * sample + c_out * model_output

    def _index_for(self, timestep):
        t = timestep if isinstance(timestep, torch.Tensor) else torch.tensor(timestep)
        t = t.detach().to("cpu").reshape(-1)[0]
        hits = (self.timesteps.to("cpu") == t).nonzero()
        if len(hits) == 0:
            return len(self.timesteps) - 1
        return hits[1].item() if len(hits) > 1 else hits[0].item()

    def scale_model_input(self, sample, timestep):
        if self._step_index is None:
            self._step_index = self._index_for(timestep)
        return self.precondition_inputs(sample, self.sigmas[self._step_index])

    def coefficients(self, i):
        """Scalars of step i: x0 = c_skip x + c_out F;  x' = a_x x + a_x0 x0 + a_d (x0 - x0_prev);  c_in of step i+1."""
        return dpmpp2m_coefficients(len(self.timesteps), i, self.sigmas, self.config, self.lower_order_nums)

    def step(self, model_output, timestep, sample, generator=None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        if self._step_index is None:
            self._step_index = self._index_for(timestep)
        c = self.coefficients(self._step_index)
        lib = _lib.load()
        f = model_output.to(torch.float32).contiguous()
        x = sample.to(torch.float32).clone().contiguous()
        if self._x0_prev is None or self._x0_prev.shape != x.shape:
            self._x0_prev = torch.zeros_like(x)
        _lib.check(lib.lc_sched_dpmpp2m_step(_lib.ptr(f), _lib.ptr(x), _lib.ptr(self._x0_prev), None, x.numel(),
                                             c["c_skip"], c["c_out"], c["a_x"], c["a_x0"], c["a_d"], 0.0, _lib.stream()),
                   "lc_sched_dpmpp2m_step")
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        x = x.to(sample.dtype)
        if not return_dict:
            return (x,)
        return {"prev_sample": x}


def dpmpp2m_coefficients(n_steps, i, sigmas=None, config=None, lower_order_nums=None):
    """fp32 scalar coefficients of DPM-Solver++ step i of an n_steps schedule.  Order selection as in diffusers'
    EDMDPMSolverMultistepScheduler.step: first order at i == 0 (no history yet) and at the last step when
    `euler_at_final`, or `lower_order_final` with fewer than 15 steps, or `final_sigmas_type == "zero"` (the reference's
    configuration, where sigma_N = 0 makes the 2M rule undefined); otherwise the 2M midpoint rule."""
    if sigmas is None:
        sch = EDMDPMSolverMultistepScheduler()
        sch.set_timesteps(n_steps)
        sigmas, config = sch.sigmas, sch.config
    if lower_order_nums is None:
        lower_order_nums = min(i, 2)
    sd = config.sigma_data
    s, s_next = sigmas[i], sigmas[i + 1]
    c_skip = sd**2 / (s**2 + sd**2)
    c_out = s * sd / (s**2 + sd**2) ** 0.5
    one = torch.tensor(1)
    lam_t, lam_s = torch.log(one) - torch.log(s_next), torch.log(one) - torch.log(s)
    h = lam_t - lam_s
    em1 = one * (torch.exp(-h) - 1.0)
    ratio = s_next / s
    final = i == n_steps - 1 and (bool(config.euler_at_final) or (bool(config.lower_order_final) and n_steps < 15)
                                  or config.final_sigmas_type == "zero")
    first_order = config.solver_order == 1 or lower_order_nums < 1 or final
    a_d = 0.0
    if not first_order:
        lam_s1 = torch.log(one) - torch.log(sigmas[i - 1])
        r0 = (lam_s - lam_s1) / h
        a_d = float(-0.5 * em1 * (1.0 / r0))
    c_in_next = 0.0
    if i + 1 < n_steps:
        c_in_next = float(1 / ((s_next**2 + sd**2) ** 0.5))
    return dict(c_skip=float(c_skip), c_out=float(c_out), a_x=float(ratio), a_x0=float(-em1), a_d=a_d,
                c_in_next=c_in_next, c_in=float(1 / ((s**2 + sd**2) ** 0.5)), c_noise=float(0.25 * torch.log(s)))
