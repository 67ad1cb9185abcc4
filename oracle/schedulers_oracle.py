"""TEST INFRASTRUCTURE — restatement of the two third-party schedulers the reference drives.

PARITY UNPINNED: the scheduler sources are not under /root/reference (pinned dependency
`diffusers==0.18.2`, environment.yaml:15, not installed here and not fetchable). The arithmetic below
restates the published algorithms of that version (`schedulers/scheduling_pndm.py`,
`schedulers/scheduling_euler_discrete.py`) and is anchored on the reference's call sites:
  SD1.5  PNDMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
         num_train_timesteps=1000, skip_prk_steps=True, steps_offset=1)   models/region_diffusion.py:35-37
         .set_timesteps / .timesteps / .step / .alphas_cumprod              :95,99,139,147,177
  SDXL   EulerDiscreteScheduler.from_pretrained(<sdxl>/scheduler)          models/region_diffusion_sdxl.py:120
         (HF config: scaled_linear 0.00085-0.012, 1000 steps, steps_offset=1, timestep_spacing="leading",
          interpolation_type="linear", prediction_type="epsilon")
         .set_timesteps / .init_noise_sigma / .scale_model_input / .step / .alphas_cumprod   :735,531,784,837,845,956
The same restatement is used on both arms when the reference loop is driven through oracle/ref_shim.py,
so what the golden vectors pin is the reference's loop logic, not this file.
"""
import numpy as np
import torch


def _scaled_linear_alphas_cumprod(beta_start, beta_end, n):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class _Out(dict):
    def __getattr__(self, k):
        return self[k]

    def __getitem__(self, k):
        if isinstance(k, int):
            return list(self.values())[k]
        return dict.__getitem__(self, k)


class PNDMSchedulerOracle:
    """PLMS branch of PNDM (skip_prk_steps=True)."""
    order = 1

    def __init__(self, beta_start=0.00085, beta_end=0.012, num_train_timesteps=1000, steps_offset=1):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.alphas_cumprod = _scaled_linear_alphas_cumprod(beta_start, beta_end, num_train_timesteps)
        self.final_alpha_cumprod = self.alphas_cumprod[0]  # set_alpha_to_one=False
        self.init_noise_sigma = 1.0
        self.ets = []
        self.counter = 0
        self.cur_sample = None
        self.timesteps = None

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round() + self.steps_offset
        # skip_prk_steps: the second-to-last timestep is repeated (N+1 model evaluations for N steps)
        plms = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        if device is not None:
            self.timesteps = self.timesteps.to(device)
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def scale_model_input(self, sample, t=None):
        return sample

    def step(self, model_output, timestep, sample, **kw):
        timestep = int(timestep)
        ratio = self.num_train_timesteps // self.num_inference_steps
        prev_timestep = timestep - ratio
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev_timestep = timestep
            timestep = timestep + ratio
        if len(self.ets) == 1 and self.counter == 0:
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            model_output = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            model_output = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            model_output = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            model_output = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        prev_sample = self._get_prev_sample(sample, timestep, prev_timestep, model_output)
        self.counter += 1
        return _Out(prev_sample=prev_sample)

    def _get_prev_sample(self, sample, timestep, prev_timestep, model_output):
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        sample_coeff = (a_prev / a_t) ** 0.5
        denom = a_t * b_prev ** 0.5 + (a_t * b_t * a_prev) ** 0.5
        return sample_coeff * sample - (a_prev - a_t) * model_output / denom


class EulerDiscreteSchedulerOracle:
    order = 1

    def __init__(self, beta_start=0.00085, beta_end=0.012, num_train_timesteps=1000, steps_offset=1):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.alphas_cumprod = _scaled_linear_alphas_cumprod(beta_start, beta_end, num_train_timesteps)
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        self.sigmas = torch.from_numpy(np.concatenate([sig[::-1], [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy())

    @property
    def init_noise_sigma(self):
        # timestep_spacing == "leading"
        return (self.sigmas.max() ** 2 + 1) ** 0.5

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(float)
        ts += self.steps_offset
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        if device is not None:
            self.sigmas = self.sigmas.to(device)
            self.timesteps = self.timesteps.to(device)

    def _index(self, timestep):
        return int((self.timesteps == timestep).nonzero()[0].item())

    def scale_model_input(self, sample, timestep):
        sigma = self.sigmas[self._index(timestep)]
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, timestep, sample, generator=None, **kw):
        i = self._index(timestep)
        sigma = self.sigmas[i]
        # diffusers 0.18.2 draws the churn noise even when s_churn == 0 (RNG side effect only)
        torch.randn(model_output.shape, dtype=model_output.dtype, generator=generator)
        sigma_hat = sigma  # gamma = 0
        pred_original_sample = sample - sigma_hat * model_output
        derivative = (sample - pred_original_sample) / sigma_hat
        dt = self.sigmas[i + 1] - sigma_hat
        return _Out(prev_sample=sample + derivative * dt, pred_original_sample=pred_original_sample)
