"""RCNetModel, host side (mirror of the reference's ``RCNet/rcnet_model.py:7-265``).

Same constructor arguments, attributes (``encoder``, ``decoder``, ``input_patch_size_image``,
``device``) and methods, so ``rcnet_main.train/validate``, ``run_rcnet_zju.py`` and existing
checkpoints work unchanged.  ``forward`` (rcnet_model.py:101-123) is the A0 dispatcher of the
hot path: encoder (gather + LoFTR kernels) -> torch decoder -> logits / sigmoid.
TensorBoard ``log_summary`` (rcnet_model.py:267-476) belongs to the training driver and is out
of scope.
"""
import torch

from . import networks


class RCNetModel(object):
    def __init__(self, input_channels_image, input_channels_depth, input_patch_size_image,
                 encoder_type, n_filters_encoder_image, n_neurons_encoder_depth,
                 decoder_type, n_filters_decoder,
                 weight_initializer='kaiming_uniform', activation_func='leaky_relu',
                 device=torch.device('cuda')):
        self.input_patch_size_image = input_patch_size_image
        self.device = device
        height, width = input_patch_size_image
        latent_size_depth = int(height // 32.0) * int(width // 32.0) * n_neurons_encoder_depth[-1]

        if 'rcnet' not in encoder_type:
            raise ValueError('Encoder type {} not supported.'.format(encoder_type))
        self.encoder = networks.RCNetEncoder(
            input_channels_image=input_channels_image,
            input_channels_depth=input_channels_depth,
            input_patch_size_image=input_patch_size_image,
            n_filters_encoder_image=n_filters_encoder_image,
            n_neurons_encoder_depth=n_neurons_encoder_depth,
            latent_size_depth=latent_size_depth,
            weight_initializer=weight_initializer,
            activation_func=activation_func,
            use_batch_norm='batch_norm' in encoder_type)

        n_skips = list(n_filters_encoder_image[:-1])[::-1] + [0]
        latent_channels = n_filters_encoder_image[-1] + n_neurons_encoder_depth[-1]
        if 'multiscale' not in decoder_type:
            raise ValueError('Decoder type {} not supported.'.format(decoder_type))
        self.decoder = networks.MultiScaleDecoder(
            input_channels=latent_channels, output_channels=1, n_resolution=1,
            n_filters=n_filters_decoder, n_skips=n_skips,
            weight_initializer=weight_initializer, activation_func=activation_func,
            output_func='linear', use_batch_norm='batch_norm' in decoder_type, deconv_type='up')
        self.to(self.device)

    # -- A0 ---------------------------------------------------------------------------------
    def forward(self, image, point, bounding_boxes, return_logits=True):
        """image N x 3 x H x W, point (sum K) x 3, bounding_boxes list[N] of K x 4
        -> (sum K) x 1 x ph x pw logits (or sigmoid)."""
        latent, skips = self.encoder(image, point, bounding_boxes)
        logits = self.decoder(x=latent, skips=skips, shape=self.input_patch_size_image)[-1]
        return logits if return_logits else torch.sigmoid(logits)

    def compute_loss(self, logits, ground_truth, validity_map, w_positive_class=1.0):
        """masked BCE-with-logits, rcnet_model.py:125-163."""
        w = torch.tensor(w_positive_class, device=logits.device)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(
            input=logits, target=ground_truth, reduction='none', pos_weight=w)
        loss = torch.sum(validity_map * loss) / torch.sum(validity_map)
        return loss, {'loss': loss}

    def parameters(self):
        return list(self.encoder.parameters()) + list(self.decoder.parameters())

    def train(self):
        self.encoder.train()
        self.decoder.train()

    def eval(self):
        self.encoder.eval()
        self.decoder.eval()

    def to(self, device):
        self.encoder.to(device)
        self.decoder.to(device)

    def channels_last(self):
        """B200 extension: keep the CNN activations NHWC so the gather kernel takes its TMA path
        and cuDNN picks tensor-core kernels.  Pure layout change - results are unchanged."""
        self.encoder.to(memory_format=torch.channels_last)
        self.decoder.to(memory_format=torch.channels_last)
        return self

    def optimize_for_inference(self):
        """B200 extension, inference only: channels_last + BatchNorm folded into the convolutions.
        Results stay within the 1e-4 tolerance on confidences (float32 re-association only)."""
        self.eval()
        self._optimized = True          # BatchNorm folded away, 4-channel stem: the state_dict no longer matches the reference's
        for net in (self.encoder, self.decoder):
            for mod in net.modules():
                if isinstance(mod, networks.Conv2d):
                    mod.fold_batch_norm_()
        self.channels_last()
        for mod in self.decoder.modules():
            if isinstance(mod, networks.DecoderBlock):
                mod.fold_upsample_()
        # the 7x7 stem reads a 3-channel image: channels-last with C = 3 has no tensor-core convolution in cuDNN (a direct
        # fp32 kernel, 2.4 ms per 16 frames); a zero fourth channel (zero weights) makes it an implicit GEMM
        enc = getattr(self.encoder, 'module', self.encoder).encoder_image
        c1 = enc.conv1.conv
        if c1.in_channels == 3 and c1.groups == 1:
            c4 = torch.nn.Conv2d(4, c1.out_channels, c1.kernel_size, c1.stride, c1.padding, c1.dilation, 1, c1.bias is not None)
            c4 = c4.to(device=c1.weight.device, dtype=c1.weight.dtype)
            with torch.no_grad():
                c4.weight.zero_()
                c4.weight[:, :3].copy_(c1.weight)
                if c1.bias is not None:
                    c4.bias.copy_(c1.bias)
            enc.conv1.conv = c4.to(memory_format=torch.channels_last).eval()
            enc.stem_pad_channels = 1
        return self

    def log_summary(self, summary_writer, tag, step, image=None, output_response=None, output_label=None,
                    output_depth=None, validity_map=None, ground_truth_label=None, ground_truth_depth=None,
                    scalars={}, n_display=4):
        """Same signature as rcnet_model.py:267-476 so ``rcnet_main.train`` (:398) and ``validate`` (:634) run unchanged on
        the mirror.  Scalars and value histograms are logged; the colourised image panels of the reference need matplotlib
        colour maps (utils/log_utils.py, SURVEY 2.1: logging is out of scope) and are skipped."""
        if summary_writer is None:
            return
        with torch.no_grad():
            for name, t in (('output_response', output_response), ('output_depth', output_depth),
                            ('ground_truth_depth', ground_truth_depth)):
                if t is not None and hasattr(summary_writer, 'add_histogram'):
                    summary_writer.add_histogram('{}_{}_distro'.format(tag, name), t[0:n_display].detach().float().cpu(),
                                                 global_step=step)
            for name, value in scalars.items():
                summary_writer.add_scalar('{}_{}'.format(tag, name), float(value), global_step=step)

    def _refuse_if_optimized(self, what):
        if getattr(self, '_optimized', False):
            raise RuntimeError("RCNetModel.%s after optimize_for_inference(): the BatchNorm layers are folded into the "
                               "convolutions and the stem has 4 input channels, so the state_dict does not have the "
                               "reference's keys any more - %s BEFORE calling optimize_for_inference()" %
                               (what, "restore the checkpoint" if what == "restore_model" else "save the model"))

    def save_model(self, checkpoint_path, step, optimizer):
        self._refuse_if_optimized("save_model")
        torch.save({
            'train_step': step,
            'radarnet_optimizer_state_dict': optimizer.state_dict(),
            'radarnet_encoder_state_dict': self.encoder.state_dict(),
            'radarnet_decoder_state_dict': self.decoder.state_dict(),
        }, checkpoint_path)

    def restore_model(self, checkpoint_path, optimizer=None):
        self._refuse_if_optimized("restore_model")
        checkpoint = torch.load(checkpoint_path, map_location=self.device)
        self.encoder.load_state_dict(checkpoint['radarnet_encoder_state_dict'])
        self.decoder.load_state_dict(checkpoint['radarnet_decoder_state_dict'])
        if optimizer is not None:
            optimizer.load_state_dict(checkpoint['radarnet_optimizer_state_dict'])
        return checkpoint['train_step'], optimizer

    def data_parallel(self):
        """The reference wraps encoder/decoder in nn.DataParallel (rcnet_model.py:259-265), which
        only yields the 'module.' key prefix on a single GPU.  Here scaling is one process per GPU
        with frame sharding (riders_b200.dist); the wrapper is kept for checkpoint-key parity."""
        self.encoder = torch.nn.DataParallel(self.encoder, device_ids=[torch.cuda.current_device()]
                                             if torch.cuda.is_available() else None)
        self.decoder = torch.nn.DataParallel(self.decoder, device_ids=[torch.cuda.current_device()]
                                             if torch.cuda.is_available() else None)
