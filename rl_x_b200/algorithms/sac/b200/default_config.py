from rl_x_b200.config_dict import ConfigDict


def get_config(algorithm_name):
    """Same keys and defaults as rl_x/algorithms/sac/pytorch/default_config.py:4-28 (bf16 autocast off: this is the fp32 parity path;
    compile_mode accepted and ignored)."""
    config = ConfigDict()

    config.name = algorithm_name

    config.device = "gpu"
    config.compile_mode = "reduce-overhead"
    config.bf16_mixed_precision_training = False
    config.total_timesteps = 1e9
    config.learning_rate = 3e-4
    config.anneal_learning_rate = False
    config.buffer_size = 1e6
    config.learning_starts = 5000
    config.batch_size = 256
    config.tau = 0.005
    config.gamma = 0.99
    config.target_entropy = "auto"
    config.log_std_min = -20
    config.log_std_max = 2
    config.nr_hidden_units = 256
    config.logging_frequency = 300
    config.evaluation_frequency = -1
    config.evaluation_episodes = 10

    config.use_cuda_graph = True   # replay the whole update (noise + ~58 kernels) as one CUDA graph
    config.gemm_engine = "auto"  # auto | simt (fp32 FFMA) | tcgen05 (3xTF32 tensor cores)

    return config
