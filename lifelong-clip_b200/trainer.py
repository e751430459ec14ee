"""Trainer mirror of the reference's methods/adapter_clip.AdapterCLIP (on methods/_trainer._Trainer)
for the lora-clip method: same online_step / online_train / online_before_task / online_after_task /
online_evaluate interface and class bookkeeping, with the hot loop routed through the fused
tower + head kernels and sharded data-parallel over one process per GPU.

What changed under the same interface:
  * label remap: the O(N*C) Python `list.index` loop (methods/adapter_clip.py:75-76) is a device
    LUT gather (llc_label_remap), bit-exact;
  * per-step tokenisation + text tower (set_token, :84) is a gather index into cached features;
  * forward/backward/AdamW are llc_* calls; loss and accuracy come back in ONE 2-float read
    (the reference syncs twice, :103-104);
  * nn.DataParallel (methods/_trainer.py:167-168: ~600 MB replicate per step) is replaced by
    persistent replicas, a rank::world shard of the combined stream+replay batch and ONE NCCL
    all-reduce of the flat LoRA gradient buffer (+2 floats for loss / #correct).
The replay memory (utils/memory.py) and the Si-Blurry sampler (utils/online_sampler.py) are used
unchanged: pass the reference's objects in.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import dp, ops
from .adapter_clip import AdapterCLIP
from .engine import FlatAdamW
from .transform import GpuTransform


class LoRAClipTrainer:
    def __init__(self, model: AdapterCLIP, class_names, *, n_classes=None, n_tasks=5, lr=1e-3,
                 online_iter=1, visible_classes='batch', memory=None, memory_provider=None,
                 memory_batchsize=0, memory_size=0, train_transform=None, test_transform=None,
                 use_amp=True, topk=1, double_softmax=True, device=None, rank=None,
                 world_size=None, sharded_input=False, use_cuda_graph=True, opt_name='adamw',
                 sched_name='default'):
        self.custom_clip = model
        self.model = model
        self.device = device or model.model.visual.proj.device
        self.all_classnames = list(class_names)
        self.n_classes = n_classes or len(class_names)
        self.n_tasks = n_tasks
        self.lr, self.online_iter, self.topk = lr, online_iter, topk
        # utils/train_utils.py:16-59. 'adamw' (weight decay 1e-5, what every script sets) and
        # 'adam' (the same update with weight decay 0) run on llc_adamw; the schedules the scripts
        # use ('default', 'const') keep the learning rate constant
        if opt_name not in ('adamw', 'adam'):
            raise NotImplementedError(f"opt_name={opt_name!r}: the fused optimizer implements "
                                      "adamw / adam (scripts/*.sh use adamw)")
        if sched_name not in ('default', 'const'):
            raise NotImplementedError(f"sched_name={sched_name!r}: only the constant schedules "
                                      "('default', 'const') of the scripts are implemented")
        self.opt_name, self.sched_name = opt_name, sched_name
        self.weight_decay = 1e-5 if opt_name == 'adamw' else 0.0
        self.visible_classes = visible_classes
        self.memory, self.memory_provider = memory, memory_provider
        self.memory_batchsize, self.memory_size = memory_batchsize, memory_size
        # train_transform / test_transform: any callable on the device batch (the reference's
        # torchvision Compose), or a GpuTransform - then the raw batch is transformed inside the
        # tower's first kernel (llc_vit_forward_tx)
        self.train_transform = train_transform or (lambda x: x)
        self.test_transform = test_transform or (lambda x: x)
        self.use_amp = use_amp          # bf16 tensor-core operands are always on; no GradScaler
        self.double_softmax = double_softmax
        self.exposed_classes, self.exposed_classes_names = [], []
        self.batch_exposed_classes, self.batch_exposed_classes_names = [], []
        self._total_classes = 0         # set by the driver loop (methods/_trainer.py:322)
        self._known_classes = 0
        ddp = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank() if ddp else 0)
        self.world = world_size if world_size is not None else (dist.get_world_size() if ddp else 1)
        # sharded_input=True: the loader already yields this rank's shard (DistributedSampler
        # style; utils/online_sampler.py:33-48 has the num_replicas/rank plumbing); False: every
        # rank sees the global batch (the reference's single DataLoader) and slices rank::world.
        self.sharded_input = sharded_input
        # forward + head + backward (~330 launches) are captured once per (batch, class list) and
        # replayed: the step is then 6 host operations instead of ~350 (matters at 32 images / GPU)
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._graph_key = None
        self.graph_kernels = 0
        self._graph_hits = 0
        self._graph_churn = 0
        self.optimizer = None
        self._lut = torch.full((self.n_classes,), -1, dtype=torch.int64, device=self.device)
        self._lut_src = None
        self._comm_stream = None

    @property
    def _scal(self):
        """(loss_sum, n_correct) of the step: the two floats behind the image tower's flat LoRA
        gradient, so that they travel in the same all-reduce."""
        if self.block_mode:
            return self._block_scal
        m = self.custom_clip.model
        return (m.visual.engine() if self.image_trainable else m.text_engine()).scal

    @property
    def block_mode(self) -> bool:
        """--method adapter-clip: the towers run block by block under autograd (adapters are
        module-owned tensors); lora-clip runs the fused tower calls."""
        return getattr(self.custom_clip, "peft_method", "lora") == "adapter"

    @property
    def text_trainable(self) -> bool:
        return self.custom_clip.peft_encoder in ('both', 'text')

    @property
    def image_trainable(self) -> bool:
        return self.custom_clip.peft_encoder in ('both', 'image')

    # ---- class bookkeeping (methods/_trainer.py:404-416, methods/adapter_clip.py:256-283) ------
    def add_new_class(self, class_name):
        # same lists as methods/_trainer.py:404-416; membership through a set (the reference's
        # `label not in list` is O(classes) per label: 0.1 ms per step at 100 classes, a visible
        # slice of a 4 ms step)
        seen = getattr(self, "_exposed_set", None)
        if seen is None or len(seen) != len(self.exposed_classes):
            seen = self._exposed_set = set(self.exposed_classes)
        grew = False
        for label in class_name.tolist():
            if label not in seen:
                seen.add(label)
                self.exposed_classes.append(label)
                grew = True
        if self.memory is not None:
            self.memory.add_new_class(cls_list=self.exposed_classes)
        if grew or len(self.exposed_classes_names) != len(self.exposed_classes):
            self.exposed_classes_names = [self.all_classnames[i] for i in self.exposed_classes]
        self.batch_exposed_classes, self.batch_exposed_classes_names = [], []
        if self.memory_size > 0:
            self.batch_exposed_classes = self.exposed_classes
            self.batch_exposed_classes_names = self.exposed_classes_names
        else:
            self.add_new_batch_class(class_name)

    def add_new_batch_class(self, class_name):
        """methods/adapter_clip.py:263-271: the classes of THIS batch join the batch-visible
        list (emptied by add_new_class when there is no replay memory)."""
        for label in class_name.tolist():
            if label not in self.batch_exposed_classes:
                self.batch_exposed_classes.append(label)
        self.batch_exposed_classes_names = [self.all_classnames[i]
                                            for i in self.batch_exposed_classes]

    def update_schedule(self, reset=False):
        """methods/adapter_clip.py:249-256 with the 'default' schedule of the scripts
        (utils/train_utils.py:34-45: a constant learning rate): reset restores self.lr."""
        if reset and self.optimizer is not None:
            self.optimizer.lr = self.lr

    def extract_vector(self, image):
        """methods/adapter_clip.py:109-112: the model wrapper's (normalised) image features."""
        with torch.no_grad():
            return self.custom_clip.encode_image(image.to(self.device))

    def report_training(self, epoch, sample_num, train_loss, train_acc):
        """methods/adapter_clip.py:285-293 (the running-time / ETA fields belong to the driver
        loop's clock and are left out)."""
        import logging
        lr = self.optimizer.lr if self.optimizer is not None else self.lr
        logging.info(f"Train | epoch:{epoch}, Sample # {sample_num} | train_loss {train_loss:.4f} "
                     f"| train_acc {train_acc:.4f} | lr {lr:.6f} | "
                     f"Num_Classes {len(self.exposed_classes)} | "
                     f"Num_Batch_Classes {len(self.batch_exposed_classes)}")

    def _class_lut(self, class_list):
        """Device LUT class id -> position in class_list; rebuilt only when the list changes."""
        key = tuple(class_list)
        if key != self._lut_src:
            lut = torch.full((self.n_classes,), -1, dtype=torch.int64)
            lut[torch.tensor(class_list, dtype=torch.int64)] = torch.arange(len(class_list))
            self._lut.copy_(lut, non_blocking=True)
            self._lut_src = key
        return self._lut

    # ---- task hooks ----------------------------------------------------------------------------
    def online_before_task(self, task_id):
        """methods/adapter_clip.py:115-127: freeze everything but LoRA, rebuild the optimizer."""
        for k, v in self.custom_clip.named_parameters():
            if "adaptmlp" not in k and "lora" not in k:
                v.requires_grad = False
        self.reset_opt()

    def _towers(self):
        m = self.custom_clip.model
        return ([m.visual] if self.image_trainable else []) + \
            ([m.text_side()] if self.text_trainable else [])

    def reset_opt(self):
        """utils/train_utils.py:27-28: AdamW(lr, weight_decay=1e-5) over the trainable tensors."""
        if not (self.image_trainable or self.text_trainable):
            raise RuntimeError("peft_encoder='none': nothing is trainable (zero-shot evaluation "
                               "only)")
        if not self.image_trainable:
            # the frozen image tower runs in its evaluation arena, whose address is not part of
            # the graph key: launch eagerly (the step is dominated by the image forward anyway)
            self.use_cuda_graph = False
        if self.block_mode:
            from .engine import ParamAdamW
            m = self.custom_clip
            self.optimizer = ParamAdamW([p for _, p in m.named_parameters()], lr=self.lr,
                                        weight_decay=self.weight_decay,
                                        on_step=m.invalidate_adapters)
            m.invalidate_adapters()         # the parameters moved into the flat buffer
            self._block_scal = torch.zeros(2, device=self.device)
            return
        self.optimizer = FlatAdamW(self._towers(), lr=self.lr, weight_decay=self.weight_decay)

    def online_after_task(self, task_id):
        """methods/adapter_clip.py:129-130: set_token(self.all_classnames[:self._total_classes]),
        so that an evaluation prediction IS the raw class id (online_evaluate compares it with the
        dataset label). `_total_classes` is maintained by the driver loop
        (methods/_trainer.py:322,355); a caller that never sets it gets every class id up to the
        largest one exposed so far."""
        total = self._total_classes
        if not total:
            total = (max(self.exposed_classes) + 1) if self.exposed_classes else 0
        self.custom_clip.set_token(self.all_classnames[:total])

    # ---- hot loop ------------------------------------------------------------------------------
    def online_step(self, images, labels, idx=None):
        """methods/adapter_clip.py:34-47."""
        seen = labels
        self._global_stream = None
        if self.sharded_input and self.world > 1:
            # class bookkeeping must see the GLOBAL batch's labels; the same
            # gather tells every rank the global stream batch size
            seen = dp.gather_labels(labels, self.world, self.device)
            self._global_stream = int(seen.numel())
        self.add_new_class(seen)
        self.custom_clip.update_class_names(self.exposed_classes_names)
        _loss, _acc, _iter = 0.0, 0.0, 0
        for _ in range(int(self.online_iter)):
            loss, acc = self.online_train([images, labels])  # nothing below writes in place
            _loss += loss
            _acc += acc
            _iter += 1
        return _loss / _iter, _acc / _iter

    def prepare_batch(self, data):
        """Host-side half of online_train (methods/adapter_clip.py:53-76): pick the visible class
        list, concatenate the replay batch, let unseen replay classes join the list. Returns
        (x, y_global [host int64], train_class_list, train_class_name_list)."""
        if self.visible_classes == 'batch':
            train_class_list = self.batch_exposed_classes
            train_class_name_list = self.batch_exposed_classes_names
        else:
            train_class_list = self.exposed_classes
            train_class_name_list = self.exposed_classes_names
        x, y = data
        if self.memory is not None and len(self.memory) > 0 and self.memory_batchsize > 0:
            memory_images, memory_labels = next(self.memory_provider)
            for i in memory_labels.unique().tolist():
                if i not in train_class_list:
                    train_class_list.append(i)
                    train_class_name_list.append(
                        self.exposed_classes_names[self.exposed_classes.index(i)])
            # stream images may already sit on the device (DevicePrefetcher); labels stay on host
            x = torch.cat([x, memory_images.to(x.device, non_blocking=True)], dim=0)
            y = torch.cat([y.cpu(), memory_labels.cpu()], dim=0)
        return x, y, train_class_list, train_class_name_list

    def online_train(self, data):
        """methods/adapter_clip.py:49-107. Returns (loss: float, acc: float) of the GLOBAL batch."""
        if self.optimizer is None:
            self.reset_opt()
        x, y, train_class_list, train_class_name_list = self.prepare_batch(data)
        B = y.shape[0]
        # data-parallel shard of the combined stream+replay batch
        if self.world > 1:
            if self.sharded_input:
                n_replay = B - data[1].shape[0]        # replay samples concatenated on this rank
                if getattr(self, "_global_stream", None) is not None:
                    B = self._global_stream + n_replay * self.world
                else:
                    B = int(dp.global_count(B, self.world, self.device))
            else:
                x, y = dp.shard_batch(x, y, self.rank, self.world)
        x = x.to(self.device, non_blocking=True)
        y = y.to(self.device, non_blocking=True)
        y_local = ops.label_remap(y, self._class_lut(train_class_list))
        if not isinstance(self.train_transform, GpuTransform):
            x = self.train_transform(x)
        self.custom_clip.set_token(train_class_name_list)
        if self.block_mode:
            if isinstance(self.train_transform, GpuTransform):
                x = self.train_transform(x)
            loss_sum, n_correct = self.block_step(x, y_local, B)
        else:
            loss_sum, n_correct = self.fused_step(x, y_local, B)
        return loss_sum, n_correct / B

    def block_step(self, x, y_local, global_batch, sync=True):
        """adapter-clip step (methods/adapter_clip.py:84-101 with the adapter blocks of
        model.py:418-442): block-by-block forward under autograd, loss + its gradient in the head
        kernels, ONE all-reduce of the flat adapter gradient (+ loss, #correct), fused AdamW."""
        m, opt = self.custom_clip, self.optimizer
        m.train()
        opt.zero_grad()
        scal = self._block_scal
        if x.shape[0] == 0:
            opt.grad_flat.zero_()
            scal.zero_()
        else:
            _, _, pred, loss, _ = m._forward_blocks(
                x, m.text_tokens, labels=y_local, inv_batch=1.0 / global_batch,
                double_softmax=self.double_softmax)
            loss.backward()
            opt.gather_grads()
            scal[0] = loss.detach()
            scal[1] = (pred == y_local).sum()
        dp.allreduce_step([opt.grad_flat], scal, self.world)
        opt.step()
        if not sync:
            return scal
        loss_sum, n_correct = scal.tolist()
        return loss_sum, n_correct

    def model_forward(self, x, y):
        """(logit, loss) as the classic methods' model_forward (methods/er_baseline.py:132-147):
        `logit` are the class probabilities the reference's criterion is applied to, `loss` the
        mean loss; both stay on the device, no parameter is updated."""
        m = self.custom_clip
        if self.block_mode:
            with torch.no_grad():
                y = y.to(self.device)
                probs, _, _, loss, _ = m._forward_blocks(
                    self.test_transform(x.to(self.device)), m.text_tokens, labels=y,
                    inv_batch=1.0 / y.shape[0], double_softmax=self.double_softmax)
            return probs, loss
        with torch.no_grad():
            eng = m.model.visual.engine()
            eng.forward(self.test_transform(x.to(self.device)), training=False)
            head = self._image_head(eng, labels=y.to(self.device), inv_batch=1.0 / y.shape[0])
        return head.probs, head.loss_rows.sum()

    # ---- the fused step ------------------------------------------------------------------------
    def _image_head(self, eng, labels=None, inv_batch=None, text=None, want_dlogits=False):
        m = self.custom_clip
        if text is not None:
            return eng.head(text, m.model.logit_scale_exp(), add_mask=m._add_mask, labels=labels,
                            double_softmax=self.double_softmax, inv_batch=inv_batch,
                            want_dlogits=want_dlogits)
        return eng.head(m._text_all, m.model.logit_scale_exp(), cls_idx=m._cls_idx,
                        add_mask=m._add_mask, labels=labels, double_softmax=self.double_softmax,
                        inv_batch=inv_batch)

    def _step_body(self, x, y_local, global_batch, force_refresh=False, tx=None):
        """forward + loss + backward; leaves the LoRA grads in the engines' grad_flat and
        (loss_sum, n_correct) of this shard in self._scal. All llc_* launches on the current
        stream. tx: llc_img_transform of the raw batch x (GpuTransform path)."""
        m = self.custom_clip
        eng = m.model.visual.engine()
        thead = None
        if self.text_trainable:
            teng = m.model.text_engine()
            from .adapter_clip import eot_rows
            if getattr(self, "_eot_src", None) is not m._tokens:
                self._eot, self._eot_src = eot_rows(m._tokens), m._tokens
            thead = teng.forward(m._tokens, self._eot, training=True, force_refresh=force_refresh)
        if not self.image_trainable:
            # peft_encoder='text': frozen image tower, forward only; the head's backward runs on
            # the compact class-token rows and only its logit gradient is used
            if tx is not None:
                eng.forward(training=False, transform=tx, n=x.shape[0])
            else:
                eng.forward(x, training=False)
            head = eng.head_compact(thead.fnorm, m.model.logit_scale_exp(), add_mask=m._add_mask,
                                    labels=y_local, double_softmax=self.double_softmax,
                                    inv_batch=1.0 / global_batch, want_dlogits=True)
            head.args.skip_logit_grad = 0
            dx_cls = torch.zeros(head.N, head.D, device=self.device)
            head.keep = head.keep + (dx_cls,)
            head.backward(dx_cls)
            d_t = ops.head_dtext(head.dlogits, head.fnorm, m.model.logit_scale_exp())
            head.keep = head.keep + (d_t,)
            teng.backward(d_t)
            ops.loss_acc(head.loss_rows, head.pred, y_local, self._scal)
            return head
        if tx is not None:
            eng.forward(training=True, force_refresh=force_refresh, transform=tx, n=x.shape[0])
        else:
            eng.forward(x, training=True, force_refresh=force_refresh)
        head = self._image_head(eng, labels=y_local, inv_batch=1.0 / global_batch,
                                text=None if thead is None else thead.fnorm,
                                want_dlogits=thead is not None)
        eng.backward_from_head(head)
        if thead is not None:
            d_t = ops.head_dtext(head.dlogits, head.fnorm, m.model.logit_scale_exp())
            head.keep = head.keep + (d_t,)
            teng.backward(d_t)
        ops.loss_acc(head.loss_rows, head.pred, y_local, self._scal)
        return head

    def _graph_signature(self, x, y_local, global_batch):
        m = self.custom_clip
        sig = (tuple(x.shape), x.dtype, tuple(y_local.shape), global_batch,
               m.model.logit_scale_exp(), self.double_softmax,
               None if m._add_mask is None else m._add_mask.data_ptr(),
               m.model.visual.engine().graph_signature(),
               isinstance(self.train_transform, GpuTransform))
        if self.text_trainable:
            return sig + (m._tokens.data_ptr(), tuple(m._tokens.shape),
                          m.model.text_engine().graph_signature())
        return sig + (m._cls_idx.data_ptr(), m._cls_idx.numel(), m._text_all.data_ptr())

    def _make_tx(self, raw, dynamic, draw=None):
        if not isinstance(self.train_transform, GpuTransform):
            return None
        return self.train_transform.struct(self.train_transform._check(raw), dynamic=dynamic,
                                           draw=draw)

    def _graph_step(self, x, y_local, global_batch):
        gpu_tf = isinstance(self.train_transform, GpuTransform)
        draw = self.train_transform.draw() if gpu_tf else None   # ONE draw per step
        key = self._graph_signature(x, y_local, global_batch)
        if key != self._graph_key:
            # a capture costs tens of ms: worth it only when the (batch, class list) signature is
            # stable (visible_classes='all', one change per task). If it keeps changing
            # (visible_classes='batch'), fall back to eager launches for good.
            if self._graph is not None and self._graph_hits < 2:
                self._graph_churn += 1
                if self._graph_churn >= 4:
                    self.use_cuda_graph = False
                    self._graph = None
                    self._graph_key = None
                    return self._step_body(x, y_local, global_batch,
                                           tx=self._make_tx(x, False, draw))
            else:
                self._graph_churn = 0
            self._graph_hits = 0
            self._graph = None
            if gpu_tf:
                x = self.train_transform._check(x)
            self._gx = torch.empty_like(x)
            self._gy = torch.empty_like(y_local)
            self._gx.copy_(x); self._gy.copy_(y_local)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):      # warm-up: arena, smem attributes, allocator pool
                self._step_body(self._gx, self._gy, global_batch, force_refresh=True,
                                tx=self._make_tx(self._gx, True, draw))
            torch.cuda.current_stream(self.device).wait_stream(side)
            # the warm-up may have (re)allocated the arenas: the key is taken after it
            key = self._graph_signature(x, y_local, global_batch)
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            tx = self._make_tx(self._gx, True, draw)     # struct with the stable dyn pointer
            with torch.cuda.graph(g):
                self._ghead = self._step_body(self._gx, self._gy, global_batch,
                                              force_refresh=True, tx=tx)
            self.graph_kernels = ops.launch_count() - n0   # libllc kernel nodes per replay
            self._graph, self._graph_key = g, key
        if gpu_tf:
            x = self.train_transform._check(x)
            self.train_transform.push(draw, x.device)   # this step's crop / flip -> device buffer
        self._gx.copy_(x, non_blocking=True)
        self._gy.copy_(y_local, non_blocking=True)
        self._graph.replay()
        self._graph_hits += 1
        return self._ghead

    def fused_step(self, x, y_local, global_batch, sync=True):
        """forward + loss + backward + gradient all-reduce + AdamW, all on the device."""
        if self.optimizer is None:
            self.reset_opt()
        if x.shape[0] == 0:
            head = self._empty_shard_step()
        elif self.use_cuda_graph:
            head = self._graph_step(x, y_local, global_batch)
        else:
            head = self._step_body(x, y_local, global_batch, tx=self._make_tx(x, False))
        engines = self.optimizer.engines()
        # image tower: gradient + (loss_sum, n_correct) in one buffer; text tower: its gradient
        dp.allreduce_step([engines[0].grad_store] + [e.grad_flat for e in engines[1:]], None,
                          self.world)
        self.optimizer.step()
        self.last_head = head
        if not sync:
            return self._scal
        loss_sum, n_correct = self._scal.tolist()  # the step's only host sync
        return loss_sum, n_correct

    def _empty_shard_step(self):
        """A rank whose shard of a ragged last batch is empty contributes zeros and still joins
        the collectives (the other ranks would block in all_reduce otherwise)."""
        for e in self.optimizer.engines():
            e.grad_store.zero_()
        return None

    # ---- evaluation (methods/adapter_clip.py:132-176, methods/_trainer.py:519-534) --------------
    @torch.no_grad()
    def online_evaluate(self, test_loader, samples_cnt=None):
        """Same dictionary as the reference. Per-task counters and the confusion matrix are
        accumulated on the device (llc_eval_accum) and read back ONCE at the end; the reference
        does two .tolist() per batch and runs sklearn on the host."""
        self.custom_clip.eval()
        m = self.custom_clip
        if self.block_mode:
            return self._block_evaluate(test_loader)
        eng = m.model.visual.engine()
        Cn = self.n_classes
        cm = torch.zeros(Cn, Cn, dtype=torch.int64, device=self.device)
        counts = torch.zeros(22, dtype=torch.int64, device=self.device)
        text = None
        if self.text_trainable:   # text features of the evaluated class list, once per call
            from .adapter_clip import eot_rows
            text = m.model.text_engine().forward(m._tokens, eot_rows(m._tokens),
                                                 training=False).fnorm
        for batch in test_loader:
            x, y = batch[0], batch[1]
            x = x.to(self.device, non_blocking=True)
            y = y.to(self.device, non_blocking=True)
            if isinstance(self.test_transform, GpuTransform):
                tt = self.test_transform
                eng.forward(training=False, transform=tt.struct(tt._check(x)), n=x.shape[0])
            else:
                eng.forward(self.test_transform(x), training=False)
            if text is not None:
                head = eng.eval_head(text, m.model.logit_scale_exp(), add_mask=m._add_mask,
                                     want_probs=False)
            else:
                head = eng.eval_head(m._text_all, m.model.logit_scale_exp(), cls_idx=m._cls_idx,
                                     add_mask=m._add_mask, want_probs=False)
            ops.eval_accum(y, head.pred, self.n_tasks, Cn, cm, counts)
        return self._eval_dict(cm, counts)

    @torch.no_grad()
    def offline_evaluate(self, test_loader, classes_names):
        """methods/adapter_clip.py:178-208: top-1 accuracy over an explicit class-name list (the
        prediction index is the position in classes_names); the counters stay on the device and
        are read once."""
        m = self.custom_clip
        m.eval()
        prev = (m._cls_key, m.text_tokens)
        m.set_token(list(classes_names))
        total = torch.zeros(2, dtype=torch.int64, device=self.device)
        t_hat = text = None
        if self.block_mode:
            t_hat = m.text_features_blocks(m.text_tokens)
        elif self.text_trainable:
            from .adapter_clip import eot_rows
            text = m.model.text_engine().forward(m._tokens, eot_rows(m._tokens),
                                                 training=False).fnorm
        for batch in test_loader:
            x = batch[0].to(self.device, non_blocking=True)
            y = batch[1].to(self.device, non_blocking=True)
            if self.block_mode:
                pred = m._forward_blocks(self.test_transform(x), m.text_tokens, t_hat=t_hat)[2]
            else:
                eng = m.model.visual.engine()
                eng.forward(self.test_transform(x), training=False)
                if text is not None:
                    pred = eng.eval_head(text, m.model.logit_scale_exp(), add_mask=m._add_mask,
                                         want_probs=False).pred
                else:
                    pred = eng.eval_head(m._text_all, m.model.logit_scale_exp(),
                                         cls_idx=m._cls_idx, add_mask=m._add_mask,
                                         want_probs=False).pred
            total[0] += (pred == y).sum()
            total[1] += y.numel()
        if prev[0] is not None:
            m.set_token(list(prev[0]))
        correct, n = total.tolist()
        return correct / n

    def _block_evaluate(self, test_loader):
        m, Cn = self.custom_clip, self.n_classes
        cm = torch.zeros(Cn, Cn, dtype=torch.int64, device=self.device)
        counts = torch.zeros(22, dtype=torch.int64, device=self.device)
        t_hat = m.text_features_blocks(m.text_tokens)      # once per call
        for batch in test_loader:
            x = batch[0].to(self.device, non_blocking=True)
            y = batch[1].to(self.device, non_blocking=True)
            _, _, pred, _, _ = m._forward_blocks(self.test_transform(x), m.text_tokens,
                                                 t_hat=t_hat)
            ops.eval_accum(y, pred, self.n_tasks, Cn, cm, counts)
        return self._eval_dict(cm, counts)

    def _eval_dict(self, cm, counts):
        counts = counts.cpu()
        if int(counts[10]) or int(counts[21]):
            # methods/_trainer.py:521-527 indexes ten-element tensors with y // n_tasks
            raise IndexError(f"label // n_tasks ({self.n_tasks}) reached bin >= 10: the "
                             "reference's _interpret_pred holds ten bins")
        num_data_l = counts[:10].float()
        correct_l = counts[11:21].float()
        if self.n_tasks != 10:
            # the reference adds the ten-bin result to zeros(n_tasks) (methods/adapter_clip.py:134,
            # 152-153), which only broadcasts for n_tasks == 10; keep the first n_tasks bins
            # when they hold everything, else report all ten
            k = self.n_tasks if float(num_data_l[self.n_tasks:].sum()) == 0 else 10
            num_data_l, correct_l = num_data_l[:k], correct_l[:k]
        avg_acc = torch.sum(correct_l) / torch.sum(num_data_l)
        task_acc = (correct_l / (num_data_l + 1e-5)).numpy().tolist()
        # sklearn.metrics.confusion_matrix(label, pred_list) with labels=None: rows / columns are
        # the sorted union of the values present in either list
        cmh = cm.cpu()
        present = ((cmh.sum(0) + cmh.sum(1)) > 0).nonzero().flatten()
        cmh = cmh[present][:, present]
        return {"avg_loss": 0.0 / torch.sum(num_data_l), "avg_acc": avg_acc, "cls_acc": task_acc,
                "task_acc": task_acc, "confusion_matrix": cmh.tolist()}

    def _interpret_pred(self, y, pred):
        """methods/_trainer.py:519-534: (ret_num_data, ret_corrects), ten float bins indexed by
        y // n_tasks, computed on the device with one llc_eval_accum launch."""
        counts = torch.zeros(22, dtype=torch.int64, device=y.device)
        ops.eval_accum(y.contiguous(), pred.contiguous(), self.n_tasks, self.n_classes, None,
                       counts)
        c = counts.cpu()
        if int(c[10]) or int(c[21]):
            raise IndexError("index out of bounds for the ten bins of _interpret_pred "
                             "(methods/_trainer.py:521-527)")
        return c[:10].float(), c[11:21].float()


class DevicePrefetcher:
    """Wraps the reference's DataLoader (pin_memory=True): the host->device copy of batch i+1 is
    issued on a side stream while batch i trains, so `online_step` receives device images.
    Yields `(images_on_device, labels_on_host, idx)` - labels stay on the host because the class
    bookkeeping of online_step (methods/_trainer.py:404-416) is host-side list logic; their 2 KB
    device copy is made inside online_train.

        for images, labels, idx in DevicePrefetcher(train_dataloader, device):
            loss, acc = trainer.online_step(images, labels, idx)
    """

    def __init__(self, loader, device):
        self.loader, self.device = loader, torch.device(device)
        self.stream = torch.cuda.Stream(self.device)

    def _stage(self, batch):
        images, labels = batch[0], batch[1]
        rest = tuple(batch[2:])
        with torch.cuda.stream(self.stream):
            dev = images.to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return dev, labels, rest, ev

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            dev, labels, rest, ev = nxt
            try:
                nxt = self._stage(next(it))      # copy of the NEXT batch overlaps this step
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            dev.record_stream(torch.cuda.current_stream(self.device))
            yield (dev, labels) + rest
