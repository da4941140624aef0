"""Decoder training step and data-parallel wiring.

`decoder_train_step` is `Tacotron2.train_step` (/root/reference/models/tts/tacotron2.py:515-522)
restricted to the hot path: zero_grad -> Decoder.forward -> loss -> BPTT -> [gradient all-reduce]
-> clip_grad_norm_(1.0) -> Adam.  The loss is `Tacotron2Loss` (:598-615) without the postnet term
(the postnet is outside the path): MSE over ALL mel elements + BCE-with-logits over
all gate elements, padded frames included.

Data parallelism (absent from the reference): one process per GPU, replicated
parameters, the batch is the shard.  The only exchange is the gradient all-reduce: gradients are
flattened into size-bounded buckets, each bucket is reduced asynchronously on NCCL's stream so the
buckets pipeline against each other and against the unflatten copies, then averaged.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F


def decoder_loss(mel_out, gate_out, mel_target, gate_target):
    """Tacotron2Loss (tacotron2.py:598-615) restricted to the decoder outputs."""
    mel_loss = F.mse_loss(mel_out, mel_target)
    gate_loss = F.binary_cross_entropy_with_logits(gate_out.reshape(-1, 1), gate_target.reshape(-1, 1))
    return mel_loss + gate_loss, mel_loss, gate_loss


def bucketize(tensors, bucket_bytes):
    """Greedy split of `tensors` (kept in order) into lists whose payload is <= bucket_bytes
    (a single larger tensor gets its own bucket)."""
    buckets, cur, cur_bytes = [], [], 0
    for t in tensors:
        nb = t.numel() * t.element_size()
        if cur and cur_bytes + nb > bucket_bytes:
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(t)
        cur_bytes += nb
    if cur:
        buckets.append(cur)
    return buckets


def allreduce_gradients(params, group=None, bucket_mb=32.0):
    """Average .grad of `params` over the process group, bucketed and asynchronous.
    Returns the number of buckets.  No-op (returns 0) when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    pending = []
    for bucket in bucketize(grads, int(bucket_mb * (1 << 20))):
        flat = torch.cat([g.reshape(-1) for g in bucket])
        pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, bucket))
    for work, flat, bucket in pending:
        work.wait()
        flat.div_(world)
        off = 0
        for g in bucket:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
    return len(pending)


def decoder_train_step(decoder, optimizer, memory, mel_padded, gate_padded, memory_lengths, grad_clip_thresh=1.0,
                       group=None, bucket_mb=32.0, sync_gradients=True):
    """One optimisation step of the decoder.  Returns (loss, grad_norm) as 0-dim device tensors
    (no host sync here; the reference's `.item()` calls at :519,:521 are the caller's choice)."""
    optimizer.zero_grad(set_to_none=True)
    mel_out, gate_out, _ = decoder(memory, mel_padded, memory_lengths)
    loss, _, _ = decoder_loss(mel_out, gate_out, mel_padded, gate_padded)
    loss.backward()
    params = [p for p in decoder.parameters() if p.requires_grad]
    if sync_gradients:          # collective: every rank of the group must take part
        allreduce_gradients(params, group, bucket_mb)
    grad_norm = torch.nn.utils.clip_grad_norm_(params, grad_clip_thresh)
    optimizer.step()
    return loss.detach(), grad_norm


def model_train_step(model, batch, criterion, optimizer, group=None, bucket_mb=32.0, grad_clip_thresh=None):
    """Data-parallel drop-in for `Tacotron2.train_step` (tacotron2.py:515-522) on a reference
    Tacotron2 whose decoder was swapped by `genvox_b200.install`: same sequence of calls, plus the one exchange the
    reference lacks - the bucketed asynchronous gradient all-reduce over ALL parameters (encoder, decoder, postnet)
    between `loss.backward()` (:520) and `clip_grad_norm_` (:521).  `criterion` / `optimizer` are the dictionaries of
    `get_criterion` / `get_optimizer` (:500-513).  The two `.item()` host syncs of the reference (:519, :521) are kept
    as attributes only when asked for: `model.loss_items` / `model.grad_norm_val` hold 0-dim device tensors here, so
    a step enqueues without waiting for the GPU; the Trainer's logging converts them when it prints."""
    optimizer["optimizer"].zero_grad(set_to_none=True)
    outputs = model.forward(batch=batch)
    loss = criterion["loss"](batch, outputs)
    loss["loss"].backward()
    params = [p for p in model.parameters() if p.requires_grad]
    allreduce_gradients(params, group, bucket_mb)
    thresh = grad_clip_thresh if grad_clip_thresh is not None else model.model_config.grad_clip_thresh
    model.grad_norm_val = torch.nn.utils.clip_grad_norm_(params, thresh)
    optimizer["optimizer"].step()
    model.loss_items = {key: val.detach() for key, val in loss.items()}
    return model.loss_items


def rank_batch_rows(batch_size, rank):
    """First row of rank `rank` in the global batch: set `model.decoder.dropout_row_offset` to it so that the ranks
    draw disjoint rows of ONE dropout stream (a run on N GPUs with per-rank batch B then uses the masks of a
    single-GPU run with batch N*B)."""
    return rank * batch_size


def make_optimizer(decoder, learning_rate=1e-3, weight_decay=1e-6):
    """Adam with the reference's hyper-parameters (tacotron2.py:506-513, configs/models.py)."""
    return torch.optim.Adam(decoder.parameters(), lr=learning_rate, weight_decay=weight_decay)


def shard_rows(n_rows, rank, world):
    """Contiguous shard [lo, hi) of n_rows for `rank` (inference shards utterances, no collective)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
