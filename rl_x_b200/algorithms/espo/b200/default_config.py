from rl_x_b200.config_dict import ConfigDict


def get_config(algorithm_name):
    """Same keys and defaults as rl_x/algorithms/espo/pytorch/default_config.py:4-30; `bf16_mixed_precision_training` defaults to
    False (fp32 parity path), `compile_mode` is accepted and ignored, B200-specific keys appended."""
    config = ConfigDict()

    config.name = algorithm_name

    config.device = "gpu"  # a CUDA device is mandatory: there is no CPU fallback
    config.compile_mode = "default"
    config.bf16_mixed_precision_training = False
    config.total_timesteps = 1e9
    config.learning_rate = 3e-4
    config.anneal_learning_rate = False
    config.nr_steps = 2048
    config.max_epochs = 300
    config.minibatch_size = 64
    config.gamma = 0.99
    config.gae_lambda = 0.95
    config.max_ratio_delta = 0.25
    config.delta_calc_operator = "mean"  # mean (in-kernel) | median (torch.median over a per-row device buffer is not built: rejected)
    config.entropy_coef = 0.0
    config.critic_coef = 0.5
    config.max_grad_norm = 0.5
    config.std_dev = 1.0
    config.action_clipping_and_rescaling = True
    config.nr_hidden_units = 256
    config.evaluation_frequency = -1
    config.evaluation_episodes = 10

    # B200-specific
    config.gemm_engine = "auto"          # auto | simt (fp32 FFMA) | tcgen05 (3xTF32 tensor cores)
    config.rollout_noise = "philox"      # philox (in-kernel counter-based normals) | torch (torch.randn on device, injected)

    return config
