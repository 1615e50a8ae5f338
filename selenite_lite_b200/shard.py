"""Channel sharding across the GPUs of one box: contiguous channel ranges, one process per GPU,
NO hot-path collective. The only communication is the optional gather of decoded audio after the compute."""
import os


def shard_range(channels, rank, world):
    """Contiguous range [lo, hi) of rank `rank`: [g*C/G, (g+1)*C/G)."""
    lo = channels * rank // world
    hi = channels * (rank + 1) // world
    return lo, hi


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def gather_audio(local_out, channels, group=None, out=None):
    """Optional gather (NCCL over NVLink on GPUs, gloo on CPU): local_out is this rank's [hi-lo][frames][2] int16
    tensor; returns the full [channels][frames][2] tensor on every rank (written into `out` when given).
    Equal shards (65536 channels over 2 / 4 / 8 GPUs) go through ONE all-gather straight into the result; shards that
    differ by a channel are padded to the largest."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_range(channels, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    # neither NCCL nor gloo has a 16-bit integer type: one I/Q (L/R) frame travels as one int32
    frames32 = local_out.contiguous().view(torch.int32)
    if all(hi - lo == biggest for lo, hi in sizes):
        if out is None:
            out = torch.empty((channels,) + tuple(local_out.shape[1:]), dtype=torch.int16, device=local_out.device)
        assert out.is_contiguous() and out.dtype == torch.int16 and out.shape[0] == channels
        dist.all_gather_into_tensor(out.view(torch.int32), frames32, group=group)
        return out
    pad = torch.zeros((biggest,) + tuple(frames32.shape[1:]), dtype=torch.int32, device=local_out.device)
    pad[:frames32.shape[0]] = frames32
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    full = torch.cat([parts[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], 0).view(torch.int16)
    if out is not None:
        out.copy_(full)
        return out
    return full


def gather_spectra(local_power, channels, group=None):
    """Optional gather of the shards' power spectra (slb_rx_spectrum_device): local_power is this rank's [hi-lo][N] float32
    tensor; returns [channels][N] on every rank (NCCL all-gather over NVLink on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_range(channels, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((biggest,) + tuple(local_power.shape[1:]), dtype=local_power.dtype, device=local_power.device)
    pad[:local_power.shape[0]] = local_power
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([parts[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], 0)
