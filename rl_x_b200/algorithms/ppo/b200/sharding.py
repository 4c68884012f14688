"""Host-side index arithmetic of the env-sharded (data-parallel) PPO update (SURVEY.md §8 e).

The rollout buffer is time-major (T, N_global) and rank r owns the env columns [r*N_local, (r+1)*N_local).  The reference
shuffles the flat index range(T*N_global) (ppo.py:273-276) and slices minibatches out of that permutation (ppo.py:277-279);
to stay reference-exact every rank walks the SAME global permutation and keeps, per minibatch, the rows it owns, in
permutation order.  Pure numpy; exercised on CPU with world_size-2 gloo tests."""
import numpy as np


def global_minibatch_sizes(batch_size, minibatch_size):
    """Sizes of batch_indices[start:start+mb] for start in range(0, B, mb) — the last one may be short (ppo.py:277-279)."""
    starts = np.arange(0, batch_size, minibatch_size)
    return np.minimum(minibatch_size, batch_size - starts).astype(np.int64)


def local_rows_of_permutation(perm, minibatch_size, global_nr_envs, local_nr_envs, rank):
    """perm: int64 permutation of range(T * global_nr_envs) (flat index k = t * N_global + env).
    Returns (local_idx, counts): local flat indices (t * N_local + local_env) of the rows this rank owns, concatenated in
    minibatch order, and how many of them fall into each global minibatch."""
    perm = np.asarray(perm, dtype=np.int64)
    t, env = np.divmod(perm, global_nr_envs)
    owner = env // local_nr_envs
    mine = owner == rank
    local_idx = (t[mine] * local_nr_envs + (env[mine] - rank * local_nr_envs)).astype(np.int64)
    mb_of_pos = np.arange(perm.shape[0]) // minibatch_size
    nmb = -(-perm.shape[0] // minibatch_size)
    counts = np.bincount(mb_of_pos[mine], minlength=nmb).astype(np.int64)
    return np.ascontiguousarray(local_idx), counts


def shard_env_slice(global_nr_envs, world_size, rank):
    """Env columns owned by `rank` (equal shards; the env count must divide evenly)."""
    if global_nr_envs % world_size != 0:
        raise ValueError("global_nr_envs must be divisible by world_size")
    n = global_nr_envs // world_size
    return slice(rank * n, (rank + 1) * n)
