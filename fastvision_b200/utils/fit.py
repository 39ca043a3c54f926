"""``Fit`` -- drop-in for the caller glue of the hot path, utils/fit.py:12-105.

``_train`` is the reference's loop (utils/fit.py:47-71); with the drop-in ``Yolov3Loss`` its ``loss.backward()`` runs the
hand-written backward kernels.  ``_val`` (utils/fit.py:73-105) keeps the reference's semantics -- loss of every batch,
per-image confidence filter + NMS, per-image mAP matching, ``fetch`` at the end -- but runs the fused step: one decode kernel
(the model is asked for the raw heads only), NMS of all images in one launch, the loss branch beside it, one matcher launch per
batch and the AP integration on the device; the reference's per-image Python loop with >= 3 host syncs per image is gone.
Checkpoint writing (fastvision.utils.checkpoints.SaveModel, out of scope) is an optional ``save_fn`` callback.
"""
import numpy as np
import torch

from ..metrics import CalculateMAP
from ..pipeline import ValStep


class Fit():

    def __init__(self, model, device, optimizer, scheduler, loss, end_epoch, start_epoch=0, train_loader=None, val_loader=None,
                 test_loader=None, data_dict=None, save_fn=None, verbose=True):
        self.model = model
        self.device = device
        self.optimizer = optimizer
        self.loss = loss
        self.start_epoch = start_epoch
        self.end_epoch = end_epoch
        self.train_loader = train_loader
        self.val_loader = val_loader
        self.test_loader = test_loader
        self.scheduler = scheduler
        self.category_names = {k: v for k, v in enumerate((data_dict or {}).get('categories', []))}
        self.save_fn = save_fn
        self.verbose = verbose
        self.history = []            # (epoch, batch, loss) of _train; last validation result in self.val_result
        self.val_result = None
        self._val_step = None

    def run_epoches(self):
        """utils/fit.py:28-45: train every epoch, validate when a val_loader is given, checkpoint through ``save_fn``."""
        epoch = self.start_epoch
        while epoch < self.end_epoch:
            self._train(epoch)
            if self.val_loader:
                self._val()
            if self.save_fn is not None:
                self.save_fn({'model': self.model, 'optimizer': self.optimizer.state_dict()}, 'last.pth')
            epoch += 1
        if self.test_loader:
            self._test()

    def _to_device(self, *tensors):
        if self.device.type != 'cuda':
            return tensors
        return tuple(t.cuda(non_blocking=True) for t in tensors)

    def _train(self, epoch):
        """utils/fit.py:47-71: forward, zero_grad, loss, backward, step; the scheduler advances once per epoch."""
        assert self.train_loader, 'train_loader can not be None'
        self.model.train()
        for step_no, batch in enumerate(self.train_loader):
            images, labels = self._to_device(*batch)
            outputs = self.model(images)
            self.optimizer.zero_grad()
            loss = self.loss(outputs, labels)
            loss.backward()                                   # CUDA backward kernels (fvb_yolov3_loss_backward_f32)
            self.optimizer.step()
            scalar = loss.item()                              # the reference reads the loss every batch too (:63)
            self.history.append((epoch, step_no, scalar))
            if self.verbose:
                print("Epoch %d batch %d loss %.6f" % (epoch + 1, step_no + 1, scalar))
        self.scheduler.step()

    def _model_core(self):
        return self.model.module if isinstance(self.model, torch.nn.DataParallel) else self.model

    def _val(self):
        map_est = CalculateMAP(map_iou_values=np.linspace(0.5, 0.95, 10))
        core = self._model_core()
        if self._val_step is None:
            self._val_step = ValStep(core.anchors_per_level, core.backbone_strides_per_level, conf_thres=0.25, iou_thres=0.45,
                                     max_det=300, ratio_box=self.loss.ratio_box, ratio_conf=self.loss.ratio_conf,
                                     ratio_cls=self.loss.ratio_cls)
        step = self._val_step
        loss_value = None
        self.model.eval()
        # In eval mode the reference's Yolov3.forward returns (head_out, results) (detection/models/yolov3.py:33-54); the fused
        # step decodes the raw heads itself, so the drop-in model is told to skip its own decode (a reference model that cannot
        # be told simply has its `results` ignored).
        had = getattr(core, "decode_in_forward", None)
        if had is not None:
            core.decode_in_forward = False
        try:
            with torch.no_grad():
                for batch in self.train_loader:                            # (sic: the reference validates on train_loader, :80)
                    images, labels = self._to_device(*batch)
                    model_out = self.model(images)
                    head_out = model_out[0] if isinstance(model_out, tuple) else model_out
                    out = step(head_out, labels)
                    loss_value = out["loss"]
                    # utils/fit.py:94-101 for the whole batch: detections [cls, conf, xyxy] + pixel-unit targets + the matcher,
                    # three launches, no host sync (row counts stay on the device until fetch)
                    map_est.process_padded(out["boxes"], out["scores"], out["cls"], out["cnt"], labels, images.size(3), images.size(2))
        finally:
            if had is not None:
                core.decode_in_forward = had
        map_each_iou, map_each_cls, map_each_cls_idx = map_est.fetch()
        loss_value = float(loss_value) if loss_value is not None else float('nan')
        self.val_result = (loss_value, map_each_iou, map_each_cls, map_each_cls_idx)
        if self.verbose:
            print(f'loss : {loss_value} map : {map_each_iou.tolist()}')
        return self.val_result

    def _test(self):
        pass
