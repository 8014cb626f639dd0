"""RC-Net training step on the B200 path (BASELINE configs[3]; mirror of the inner loop of ``rcnet_main.train``,
RCNet/rcnet_main.py:296-358 with the hyper-parameters of RCNet/train_rcnet_zju.py:28-62).

  labels      |gt - z| < max_distance_correspondence where gt > 0, validity = gt > 0          rcnet_main.py:300-331
  forward     RCNetModel.forward(image, radar_point, boxes, return_logits=True)                 :342
              (backbone / decoder: torch autograd over cuDNN; gather: ops.RoiGatherFunction with the deterministic backward
              kernel; LoFTR: ops.LoftrLayerFunction = riders_loftr_layer_fwd_train / riders_loftr_layer_bwd)
  loss        masked BCE-with-logits, pos_weight 2.5                                            rcnet_model.py:125-163
  backward    loss.backward()                                                                   :358
  all-reduce  gradients averaged over the ranks in flat buckets over NCCL (one process per GPU; the reference is
              single-GPU)
  step        Adam(lr 2e-4)                                                                      :359

``bf16=True`` runs forward + loss under ``torch.autocast(bfloat16)``: cuDNN convolutions and the LoFTR tokens in bfloat16,
parameters and the optimizer state float32 (master weights), loss reduction float32.
"""
import torch

from . import dist as rdist


class TrainStep(object):
    def __init__(self, model, lr=2e-4, w_positive_class=2.5, max_distance_correspondence=0.5, bf16=True,
                 set_invalid_to_negative_class=False):
        self.model = model
        model.train()
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer = torch.optim.Adam(self.params, lr=lr)
        self.w_positive_class = float(w_positive_class)
        self.max_dist = float(max_distance_correspondence)
        self.bf16 = bool(bf16)
        self.invalid_negative = bool(set_invalid_to_negative_class)
        self.allreduce_events = []

    def labels(self, ground_truth_depth, radar_point):
        """ground_truth_depth (R,1,ph,pw) lidar crops, radar_point (R,3) -> label, validity (rcnet_main.py:296-331)."""
        z = radar_point[:, 2].view(-1, 1, 1, 1)
        near = (ground_truth_depth - z).abs() < self.max_dist
        label = (near & (ground_truth_depth > 0)).to(torch.float32)
        validity = torch.ones_like(label) if self.invalid_negative else (ground_truth_depth > 0).to(torch.float32)
        return label, validity

    def step(self, image, radar_point, bounding_boxes, ground_truth_depth, time_allreduce=False):
        """One optimisation step; returns the (detached, device) loss."""
        label, validity = self.labels(ground_truth_depth, radar_point)
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=self.bf16 and image.is_cuda):
            logits = self.model.forward(image, radar_point, bounding_boxes, return_logits=True)
            loss, _ = self.model.compute_loss(logits.float(), label, validity, w_positive_class=self.w_positive_class)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if rdist.world_size() > 1:
            if time_allreduce:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            rdist.allreduce_gradients(self.params)
            if time_allreduce:
                b.record()
                self.allreduce_events.append((a, b))
        self.optimizer.step()
        return loss.detach()
