"""Device-side schedulers used by the samplers.

The reference delegates to the third-party diffusers 0.18.2 schedulers (PNDM/PLMS for SD1.5,
models/region_diffusion.py:35-37; EulerDiscrete for SDXL, models/region_diffusion_sdxl.py:120). Those
sources are not part of the reference tree, so the published algorithms are restated here; only the
calls the reference makes are provided: set_timesteps / timesteps / scale_model_input / step /
init_noise_sigma / alphas_cumprod.  All state lives on the sampling device; `step` never synchronises.
"""
import numpy as np
import torch


def _alphas_cumprod(beta_start, beta_end, n):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class EulerDiscreteScheduler:
    """EulerDiscrete, epsilon prediction, scaled-linear betas, `leading` spacing, steps_offset=1 (SDXL config)."""
    order = 1

    def __init__(self, beta_start=0.00085, beta_end=0.012, num_train_timesteps=1000, steps_offset=1):
        self.num_train_timesteps, self.steps_offset = num_train_timesteps, steps_offset
        self.alphas_cumprod = _alphas_cumprod(beta_start, beta_end, num_train_timesteps)  # host copy (predict_x0)
        self._sig_all = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        self.sigmas_host = np.concatenate([self._sig_all[::-1], [0.0]]).astype(np.float32)
        self.timesteps_host = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy()
        self.timesteps = torch.from_numpy(self.timesteps_host)

    @property
    def init_noise_sigma(self):
        return float((self.sigmas_host.max() ** 2 + 1) ** 0.5)

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(float) + self.steps_offset
        sig = np.interp(ts, np.arange(0, len(self._sig_all)), self._sig_all)
        self.sigmas_host = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.timesteps_host = ts
        self.timesteps = torch.from_numpy(ts)  # host tensor: the loop's `t > ...` tests never touch the device

    def index_of(self, timestep):
        return int(np.nonzero(self.timesteps_host == float(timestep))[0][0])

    def sigma(self, timestep):
        return float(self.sigmas_host[self.index_of(timestep)])

    def dt(self, timestep):
        i = self.index_of(timestep)
        return float(self.sigmas_host[i + 1] - self.sigmas_host[i])

    def scale_model_input(self, sample, timestep):
        s = self.sigma(timestep)
        return sample / ((s * s + 1.0) ** 0.5)

    def step(self, model_output, timestep, sample, **kw):
        # x + eps * (sigma_next - sigma); the reference's `derivative` equals eps for epsilon prediction
        return {"prev_sample": (sample.float() + model_output.float() * self.dt(timestep)).to(sample.dtype)}


class PNDMScheduler:
    """PNDM with skip_prk_steps=True (pure PLMS), steps_offset=1 — N+1 model evaluations for N steps."""
    order = 1

    def __init__(self, beta_start=0.00085, beta_end=0.012, num_train_timesteps=1000, steps_offset=1):
        self.num_train_timesteps, self.steps_offset = num_train_timesteps, steps_offset
        self.alphas_cumprod = _alphas_cumprod(beta_start, beta_end, num_train_timesteps)
        self.final_alpha_cumprod = float(self.alphas_cumprod[0])
        self.init_noise_sigma = 1.0
        self.ets, self.counter, self.cur_sample = [], 0, None
        self.timesteps = None

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round() + self.steps_offset
        plms = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(plms)
        self.ets, self.counter, self.cur_sample = [], 0, None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, **kw):
        timestep = int(timestep)
        ratio = self.num_train_timesteps // self.num_inference_steps
        prev_timestep = timestep - ratio
        model_output = model_output.float()
        sample = sample.float()
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
        a_t = float(self.alphas_cumprod[timestep])
        a_prev = float(self.alphas_cumprod[prev_timestep]) if prev_timestep >= 0 else self.final_alpha_cumprod
        b_t, b_prev = 1 - a_t, 1 - a_prev
        coeff = (a_prev / a_t) ** 0.5
        denom = a_t * b_prev ** 0.5 + (a_t * b_t * a_prev) ** 0.5
        self.counter += 1
        return {"prev_sample": coeff * sample - (a_prev - a_t) * model_output / denom}
