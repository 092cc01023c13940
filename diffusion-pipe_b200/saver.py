"""Model / adapter export and checkpoint cadence around `train_batch` — the reference's utils/saver.py against the new
engine.

Behaviour kept from the reference (utils/saver.py:47-176):
  * a saved model is assembled by pipeline stage 0 of data-parallel replica 0 from one partial state dict per stage
    (keys = `p.original_name`), written by the replica-0 rank of every stage into `<save_dir>/tmp/` and merged after a
    barrier; adapters keep only the trainable parameters (`.default` / `.modules_to_save` stripped from the names);
    `save_dtype` casts on the way out; the run's TOML is copied next to the weights;
  * `process_epoch` / `process_step` decide when to export (`save_every_n_epochs`, `save_every_n_steps`) and when to
    checkpoint (`checkpoint_every_n_epochs`, `checkpoint_every_n_minutes` — rank 0's clock, broadcast — and the
    user's `save` / `save_quit` signal files in the run directory);
  * training state goes through `engine.save_checkpoint(..., client_state={step, examples, custom_loader})`.
What differs: Flux full-model export stays in the diffusers layout this engine trains in (`model.safetensors`); the
reference additionally re-lays it out to BFL names (models/flux.py:257-288, outside the hot path).
"""
import os
import shutil
import sys
import time

import torch

from .pipe import dist


def _main():
    return dist.get_rank() == 0


class Saver:
    def __init__(self, args, config, is_adapter, save_root, model, train_dataloader, model_engine, pipeline_model):
        self.args, self.config, self.is_adapter = args, config, is_adapter
        self.save_root = str(save_root)
        self.model, self.train_dataloader = model, train_dataloader
        self.model_engine, self.pipeline_model = model_engine, pipeline_model
        self.last_checkpoint_time = None

    # ---- export ------------------------------------------------------------------------------------------------------
    def _partial_state_dict(self, trainable_only):
        sd = {}
        for name, p in self.pipeline_model.named_parameters():
            if trainable_only and not p.requires_grad:
                continue
            on = getattr(p, 'original_name', None)
            if on is None:
                if trainable_only and _main():
                    print(f'WARNING: parameter {name} requires_grad but has no original_name; not saving it')
                continue
            if trainable_only:
                on = on.replace('.default', '').replace('.modules_to_save', '')
            t = p.detach()
            if 'save_dtype' in self.config:
                t = t.to(self.config['save_dtype'])
            sd[on] = t.to('cpu').contiguous()
        return sd

    def _export(self, name, trainable_only, write):
        grid = self.model_engine.grid
        dp_id, stage_id = grid.get_data_parallel_rank(), grid.get_pipe_parallel_rank()
        save_dir = os.path.join(self.save_root, name)
        tmp_dir = os.path.join(save_dir, 'tmp')
        if dp_id == 0 and stage_id == 0:
            os.makedirs(tmp_dir, exist_ok=False)
        dist.barrier()
        if dp_id == 0:
            torch.save(self._partial_state_dict(trainable_only), os.path.join(tmp_dir, f'state_dict_{stage_id}.bin'))
        dist.barrier()
        if dp_id == 0 and stage_id == 0:
            state_dict = {}
            for fn in sorted(os.listdir(tmp_dir)):
                if fn.endswith('.bin'):
                    state_dict.update(torch.load(os.path.join(tmp_dir, fn), map_location='cpu', weights_only=True))
            write(save_dir, state_dict)
            cfg_path = getattr(self.args, 'config', None)
            if cfg_path and os.path.isfile(cfg_path):
                shutil.copy(cfg_path, save_dir)
            shutil.rmtree(tmp_dir)
        dist.barrier()
        return save_dir

    def save_adapter(self, name):
        return self._export(name, True, self.model.save_adapter)

    def save_full_model(self, name):
        return self._export(name, False, self.model.save_model)

    def save_model(self, name):
        if _main():
            print(f'Saving model to directory {name}')
        return self.save_adapter(name) if self.is_adapter else self.save_full_model(name)

    # ---- checkpoints -------------------------------------------------------------------------------------------------
    def save_checkpoint(self, step, examples):
        self.model_engine.save_checkpoint(self.save_root, client_state={
            'step': step, 'examples': examples, 'custom_loader': self.train_dataloader.state_dict()},
            save_latest=True, exclude_frozen_parameters=True)

    def need_to_checkpoint(self, epoch=None):
        if epoch is not None:
            if 'checkpoint_every_n_epochs' in self.config and epoch % self.config['checkpoint_every_n_epochs'] == 0:
                self.last_checkpoint_time = time.time()
                return True
            return False
        if 'checkpoint_every_n_minutes' not in self.config:
            return False
        due = False
        if _main():                                   # rank 0's clock decides for everybody
            now = time.time()
            if self.last_checkpoint_time is None:
                self.last_checkpoint_time = now
            elif (now - self.last_checkpoint_time) / 60 > self.config['checkpoint_every_n_minutes']:
                due, self.last_checkpoint_time = True, now
        holder = [due]
        dist.broadcast_object_list(holder, src=0)
        return holder[0]

    def process_epoch(self, epoch, step, examples):
        """call after every step; returns (new epoch or None when training is over, checkpointed, saved)"""
        checkpointed = saved = False
        if self.train_dataloader.epoch != epoch:
            if self.need_to_checkpoint(epoch):
                self.save_checkpoint(step, examples)
                checkpointed = True
            if 'save_every_n_epochs' in self.config and epoch % self.config['save_every_n_epochs'] == 0:
                self.save_model(f'epoch{epoch}')
                saved = True
            epoch = self.train_dataloader.epoch
            if epoch > self.config.get('epochs', 1 << 30):
                return None, checkpointed, saved
            if _main():
                print(f'Started new epoch: {epoch}')
        return epoch, checkpointed, saved

    def process_step(self, step, examples):
        checkpointed = saved = False
        manual_save = manual_quit = False
        for fn, quit_too in (('save', False), ('save_quit', True)):
            path = os.path.join(self.save_root, fn)
            if os.path.isfile(path):
                manual_save, manual_quit = True, quit_too
                dist.barrier()
                if _main():
                    os.remove(path)
                break
        if 'save_every_n_steps' in self.config and step % self.config['save_every_n_steps'] == 0:
            self.save_model(f'step{step}')
            saved = True
        if self.need_to_checkpoint() or manual_save:
            self.save_checkpoint(step, examples)
            checkpointed = True
        if manual_quit:
            print('Manually quitting')
            sys.exit()
        return checkpointed, saved
