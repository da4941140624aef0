"""Drop-in wiring of the reference's callers around the B200-native decoder.

The reference objects are used UNMODIFIED; what they lack for the 8 x B200 box is added from outside, through the hooks
their own interfaces offer:

  * `enable_data_parallel(model, rank, world, ...)` - the trainer's step and data wiring
    (/root/reference/core/trainer/__init__.py:112-142, :193-236; /root/reference/models/tts/__init__.py:72-79):
      - `model.get_train_dataloader` keeps its signature but hands the `TextMelDataset` to a `DistributedSampler`
        (one disjoint shard per rank, reshuffled per epoch; the reference uses `shuffle=True`, no sampler);
      - `model.train_step` keeps its signature (`batch, criterion, optimizer`, tacotron2.py:515-522) and gains the one
        exchange step the reference has none of: the bucketed NCCL all-reduce of the gradients between `loss.backward()`
        (:520) and `clip_grad_norm_` (:521) (`genvox_b200.training.model_train_step`);
      - the ranks draw disjoint rows of ONE dropout stream (`decoder.dropout_row_offset`).
  * `rank_zero_io(trainer, rank)` - everything the Trainer writes (config dump, runner copy, W&B, eval plots, checkpoints:
    core/trainer/__init__.py:73-105, :163-191, :222-233) happens on rank 0 only.
  * `tts_batch(synthesizer, texts)` - `Synthesizer.tts` (/root/reference/core/synthesizer.py:27-45) for a LIST of texts:
    one batched decode (per-row gate stop on the device) instead of one utterance and one host sync per decoder step.

Nothing here touches the hot path itself; the decoder is `genvox_b200.Decoder`, swapped in by `genvox_b200.install`.
"""
import types

import torch

from . import synthesis
from .training import model_train_step, rank_batch_rows


def enable_data_parallel(model, rank, world, group=None, bucket_mb=32.0, seed=0):
    """Patch a reference `Tacotron2` (decoder already swapped by `genvox_b200.install`, or not: the exchange is
    decoder-agnostic) for one-process-per-GPU data parallel training.  Returns `model`."""
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler

    orig_loader = model.get_train_dataloader

    def get_train_dataloader(self, dump_dir, num_loader_workers, batch_size):
        base = orig_loader(dump_dir=dump_dir, num_loader_workers=num_loader_workers, batch_size=batch_size)
        sampler = DistributedSampler(base.dataset, num_replicas=world, rank=rank, shuffle=True, seed=seed, drop_last=False)
        if hasattr(self.decoder, "dropout_row_offset"):
            self.decoder.dropout_row_offset = rank_batch_rows(batch_size, rank)
        self.train_sampler = sampler             # the training loop calls sampler.set_epoch(epoch) through `on_epoch`
        return DataLoader(base.dataset, num_workers=num_loader_workers, batch_size=batch_size, sampler=sampler,
                          collate_fn=base.collate_fn)

    def train_step(self, batch, criterion, optimizer):
        model_train_step(self, batch, criterion, optimizer, group=group, bucket_mb=bucket_mb)

    def on_epoch(self, epoch):
        if getattr(self, "train_sampler", None) is not None:
            self.train_sampler.set_epoch(epoch)

    model.get_train_dataloader = types.MethodType(get_train_dataloader, model)
    model.train_step = types.MethodType(train_step, model)
    model.on_epoch = types.MethodType(on_epoch, model)
    model.dp_rank, model.dp_world = rank, world
    return model


def rank_zero_io(trainer, rank):
    """Keep the reference Trainer's file / network output on rank 0: the other ranks train with evaluation, W&B and
    checkpoint writing off (flags the Trainer already has, core/trainer/__init__.py:48-50, :222-233) and a private
    scratch experiment directory, so nothing they would write collides with rank 0's."""
    if rank == 0:
        return trainer
    trainer.config.run_eval = False
    trainer.config.use_wandb = False
    trainer.exp_dir = f"{trainer.exp_dir}.rank{rank}.scratch"
    return trainer


@torch.no_grad()
def tts_batch(synthesizer, texts, max_batch=64, vocode=True):
    """`Synthesizer.tts` (core/synthesizer.py:27-45) for a list of texts.  Tokenisation and vocoding are the reference's own
    (`text_processor`, `audio_processor`); the decode is ONE batched call per `max_batch` utterances.  Returns a list of
    the dictionaries `tts` returns (numpy arrays, batch dimension squeezed; "waveform" / "sampling_rate" when `vocode`)."""
    rows = []
    for text in texts:
        tokens = synthesizer.text_processor.tokenize(text)
        rows.append(synthesizer.text_processor.tokens_to_indices(tokens))
    outs = synthesis.batched_inference(synthesizer.tts_model, rows, max_batch=max_batch)
    results = []
    for o in outs:
        r = {key: val.squeeze(0).cpu().numpy() for key, val in o.items()}
        if vocode and "mel_outputs_postnet" in r:
            fs, wav = synthesizer.audio_processor.convert_mel2wav(mel=r["mel_outputs_postnet"])
            r["waveform"], r["sampling_rate"] = wav, fs
        results.append(r)
    return results
