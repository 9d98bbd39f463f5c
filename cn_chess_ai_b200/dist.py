"""Multi-GPU plumbing: one process per GPU (torch.distributed), env / replay shards per rank with no
data-path collective, and ONE exchange step per TD update: the sum of the compact gradient buffers
(173,018 FP32 = 0.69 MB) over the ranks followed by the identical SGD step on every rank.  Two ways:
  * fused (default): after connect_peers() every applied update of the library exchanges its gradient INSIDE the gradient
    contraction kernel over NVLink / NVSwitch peer memory (CUDA IPC): reduce-scatter of the 16-row blocks to their owner
    ranks, all-gather of the sums, SGD -- td_update_replay(apply=True) / td_update_replay_n / train();
  * baseline: ncclAllReduce through torch.distributed + DQN.apply_grads().
torch is plumbing here (process group, stream, handle exchange); the kernels are the library's."""
import os


def shard(n_total, rank, world):
    """contiguous block partition of n_total units: (first, count); counts differ by at most one"""
    base, rem = divmod(int(n_total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


class DeviceArray:
    """zero-copy view of a device buffer owned by libxq_b200 (for torch.as_tensor)"""

    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def grad_tensor(dqn, device):
    import torch
    ptr, n = dqn.grad_buffer()
    return torch.as_tensor(DeviceArray(ptr, n), device=device)


def allreduce_sum_(tensor, group=None):
    """in-place sum over ranks; no-op without an initialised process group"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def connect_peers(dqn, device, group=None):
    """exchange the CUDA IPC handles of the gradient exchange buffers over the process group and map the peers' buffers:
    afterwards dqn.dist_allreduce_apply() replaces all_reduce + apply_grads with one fused peer-memory kernel per rank"""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.from_numpy(dqn.dist_export()).to(device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    handles = torch.stack(gathered).cpu().numpy()
    dqn.dist_connect(rank, world, handles)
    dist.barrier(group)             # every rank has mapped every buffer before the first flag is written
    return world


class DataParallelLearner:
    """Config 4: every rank owns envs [first, first+count) (global ids keep trajectories independent of the GPU
    count), its own replay shard and a replica of the Q-network initialised from the same seed.  A training
    step = local self-play, local TD gradients at batch `batch` per GPU, all-reduce, identical update."""

    def __init__(self, total_envs, replay_capacity, batch, seed=0, eps=0.1, lr=0.001, gamma=0.99, mode=0, train_done=True):
        import torch
        from . import BatchedEnv, DQN, ReplayBuffer
        self.rank, self.world, self.local = rank_world()
        self.device = torch.device("cuda", self.local)
        first, count = shard(total_envs, self.rank, self.world)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.env = BatchedEnv(count, device=self.local, seed=seed, env_id0=first)
        self.dqn = DQN((1260, 128, 8100), lr=lr, gamma=gamma, device=self.local, seed=seed, mode=mode)
        self.replay = ReplayBuffer(shard(replay_capacity, self.rank, self.world)[1], device=self.local)
        self.env.set_stream(stream)
        self.dqn.set_stream(stream)
        self.batch, self.eps, self.lr, self.seed, self.train_done = batch, eps, lr, seed, train_done
        self.updates = 0
        self.fused = self.world > 1 and os.environ.get("XQ_DIST", "fused") != "nccl"
        if self.fused:
            connect_peers(self.dqn, self.device)
        self.grads = None if self.fused else grad_tensor(self.dqn, self.device)

    def collect(self, n_plies):
        from . import collect
        collect(self.dqn, self.env, self.replay, n_plies, self.eps, self.train_done)

    def update(self, use_target_net=True):
        from . import td_update_replay
        if self.fused:          # contraction -> exchange -> SGD in one kernel; a timed-out exchange makes the call fail (sticky)
            td_update_replay(self.dqn, self.replay, self.batch, self.seed + 1000003 * self.rank, self.updates, use_target_net, self.lr, apply=True)
        else:
            td_update_replay(self.dqn, self.replay, self.batch, self.seed + 1000003 * self.rank, self.updates, use_target_net, self.lr, apply=False)
            allreduce_sum_(self.grads)
            self.dqn.apply_grads(self.lr)
        self.updates += 1

    def replicas_identical(self):
        """every rank holds the same parameter bytes (digest all-gathered through the library's own peer-memory exchange)"""
        d = self.dqn.params_digest()
        if self.world == 1:
            return True
        if self.fused:
            allv = self.dqn.dist_allgather(d)
        else:
            import torch
            import torch.distributed as dist
            t = torch.from_numpy(d).to(self.device)
            g = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(g, t)
            allv = torch.stack(g).cpu().numpy()
        return bool((allv == allv[0]).all())
