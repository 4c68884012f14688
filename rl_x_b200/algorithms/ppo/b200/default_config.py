from rl_x_b200.config_dict import ConfigDict


def get_config(algorithm_name):
    """Same keys and defaults as rl_x/algorithms/ppo/pytorch/default_config.py:4-30.  Differences, all explicit:
    `bf16_mixed_precision_training` defaults to False (this build is the fp32 parity path), `compile_mode` is accepted and
    ignored (no tracing compiler here), and three B200-specific keys are appended."""
    config = ConfigDict()

    config.name = algorithm_name

    config.device = "gpu"  # a CUDA device is mandatory: there is no CPU fallback
    config.compile_mode = "default"
    config.bf16_mixed_precision_training = False
    config.total_timesteps = 1e9
    config.learning_rate = 3e-4
    config.anneal_learning_rate = False
    config.nr_steps = 2048
    config.nr_epochs = 10
    config.minibatch_size = 64
    config.gamma = 0.99
    config.gae_lambda = 0.95
    config.clip_range = 0.2
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
    config.exact_global_permutation = True   # multi-GPU: reference-exact global shuffle (ppo.py:273-276) vs per-rank local shuffles
    config.gradient_exchange = "peer"    # multi-GPU: peer (library all-reduce kernel over NVLink peer memory) | nccl (torch.distributed)
    config.rollout_noise = "philox"      # philox (in-kernel counter-based normals) | torch (torch.randn on device, injected)

    return config
