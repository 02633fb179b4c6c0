"""`InvarianceManifold`: learn an invariance manifold for a template neuron, then align it to other neurons.

Mirror of the reference interface laminr/inv_temp_learn_match.py:23-650: same constructor / `learn` / `match` /
`forward` / `get_images` / getter signatures, defaults, return values (NumPy arrays on the host) and error behaviour.
The loops keep the reference's control flow exactly (scheduler, stop rules, RNG draw order); what changes is what
runs inside a step:

  * simulated populations (`neuron_models.simulated*`) and conv-core + Gaussian-readout encoders (`neuron_models.FiringRateEncoder`)
    take the FUSED path (laminr_b200/fused.py): a fixed sequence
    of C-ABI launches on static buffers, CUDA-graph replayed, no autograd, no per-step host sync;
  * any other `nn.Module` response model takes the GENERIC path: the same fused INR / constraint / contrastive
    operators, composed by stock autograd around the user's model.

With `gaussian_blur_sigma=` (the GaussianBlur gradient preconditioner, featurevis/ops.py:378-423) the loops run the generic path with the blur as a
backward hook.  Out of scope here: GIF export (`save_*_as_gif`, needs matplotlib / lipstick).
"""
import numpy as np
import torch
from torch import nn
from tqdm import tqdm

from . import ops
from .contrastive_loss import SimCLROnGrid
from .fused import LearnStepper, MatchStepper, fused_response_supported
from .inr import INRTemplates
from .utils.general import SingleNeuronModel, infer_device
from .utils.img_transforms import FixEnergyNormClip, StandardizeClip
from .utils.training import ImprovementChecker, JitteringGridDatamodule, ParamReduceOnPlateau, check_activity_requirements

_BAR = "{l_bar}{bar:20}{r_bar}"


class InvarianceManifold:
    def __init__(self, response_predicting_model, meis_dict, device=None, widths=[50] * 4, positional_encoding_dim=50,
                 positional_encoding_projection_scale=10, aux_positional_encoding_dim=50, aux_positional_encoding_projection_scale=0.1,
                 periodic_invariance=True, nonlinearity=nn.Tanh, final_nonlinearity=nn.Tanh, weights_scale=0.1, bias=True,
                 required_img_norm=None, required_img_std=None, pixel_value_lower_bound=None, pixel_value_upper_bound=None,
                 precision=None, use_cuda_graph=True):
        if required_img_std is None and required_img_norm is None:
            raise ValueError("Either std or norm should be provided.")
        elif required_img_std is not None and required_img_norm is not None:
            raise ValueError("Only one of std or norm should be provided.")
        self.response_predicting_model = response_predicting_model
        channels, height, width = list(meis_dict.values())[0]["mei"].shape
        self.periodic_invariance = periodic_invariance
        if device is None:
            device = infer_device(response_predicting_model)
        if precision is None:
            # the tcgen05 path (split-bf16 x3 operands, fp32 accumulate: fp32-class results) serves the pipeline default network
            # (4 hidden layers of width 50); any other shape runs on the fp32 CUDA-core path
            precision = "bf16x3" if (len(widths) == 4 and max(widths) <= INRTemplates.KERNEL_WIDTH and positional_encoding_dim <= 64) else "fp32"
        self.precision, self.use_cuda_graph = precision, use_cuda_graph
        self.template = INRTemplates(
            img_res=(height, width), num_templates=1, out_channels=channels, widths=widths,
            positional_encoding_projection_scale=positional_encoding_projection_scale, positional_encoding_dim=positional_encoding_dim,
            aux_positional_encoding_dim=aux_positional_encoding_dim,
            aux_positional_encoding_projection_scale=aux_positional_encoding_projection_scale, periodic_invariance=periodic_invariance,
            nonlinearity=nonlinearity, final_nonlinearity=final_nonlinearity, weights_scale=weights_scale, bias=bias, precision=precision,
        ).to(device)
        self.meis_dict = meis_dict
        self.pixel_min, self.pixel_max = pixel_value_lower_bound, pixel_value_upper_bound
        mean = (self.pixel_min + self.pixel_max) / 2
        if required_img_norm is not None:
            self.img_transforms = FixEnergyNormClip(norm=required_img_norm, mean=mean, pixel_min=self.pixel_min, pixel_max=self.pixel_max).to(device)
        else:
            self.img_transforms = StandardizeClip(mean=mean, std=required_img_std, pixel_min=self.pixel_min, pixel_max=self.pixel_max).to(device)

    # ------------------------------------------------------------------------------------------------------------------
    def _fused_ok(self, gaussian_blur=None):
        return fused_response_supported(self.response_predicting_model) and gaussian_blur is None

    @staticmethod
    def _blur(sigma):
        if sigma is None:
            return None
        from .featurevis.ops import GaussianBlur
        return GaussianBlur(sigma)

    # ------------------------------------------------------------------------------------------------------------------
    # learn
    # ------------------------------------------------------------------------------------------------------------------
    def learn(self, template_neuron_idx, gaussian_blur_sigma=None, steps_per_epoch=50, grid_points_per_dim=20, neighbor_size=0.1,
              temperature=0.3, reg_decr=0.8, min_epochs=10, lr=1e-3, additional_epochs=0, reg_scale=2, requirements=None,
              num_max_epochs=1000, verbose=False, save_results_every_n_epochs=None):
        device = next(self.template.parameters()).device
        self.template_neuron_idx = template_neuron_idx
        gaussian_blur = self._blur(gaussian_blur_sigma)

        dm = JitteringGridDatamodule(num_invariances=1, grid_points_per_dim=grid_points_per_dim, steps_per_epoch=steps_per_epoch)
        dm.train_dataloader()  # the reference builds (and discards) one loader here: one jitter draw from the CPU generator (:137)
        grid = dm.grid.to(device)
        grid_reg = SimCLROnGrid(num_invariances=1, grid_points_per_dim=grid_points_per_dim, neighbor_size=neighbor_size, temperature=temperature,
                                with_periodic_invariances=self.periodic_invariance, with_round_neighbor=False).to(device)
        reg_scheduler = ParamReduceOnPlateau(factor=reg_decr, patience=5, threshold=0.005, mode="max")
        if requirements is None:
            requirements = dict(avg=0.99, std=1.0, necessary_min=0.98)
        template_mei_act = self.meis_dict[template_neuron_idx]["activation"]
        template_rf_mask = self.meis_dict[template_neuron_idx]["mask"]
        template_neuron_model = SingleNeuronModel(self.response_predicting_model, template_neuron_idx)
        fused = self._fused_ok(gaussian_blur)
        if fused:
            stepper = LearnStepper(self.template, self.img_transforms, self.response_predicting_model, template_neuron_idx, template_rf_mask,
                                   template_mei_act, grid_reg, self.pixel_min, self.pixel_max, grid_points_per_dim, lr=lr, reg_scale=reg_scale,
                                   precision=self.precision, use_graph=self.use_cuda_graph)
            self._learn_stepper = stepper
        else:
            optimizer = torch.optim.Adam(self.template.func.parameters(), lr=lr)
            mask_dev = torch.as_tensor(template_rf_mask, dtype=torch.float32).to(device)

        def snapshot():
            if fused:  # images / activations of the un-jittered grid straight from the stepper's launches (what :394-405 computes through `forward`)
                stepper.evaluate(grid)
                return stepper.z.cpu().numpy(), stepper.act[0].cpu().numpy()
            return self.get_template_images_and_activations(grid, template_neuron_model, gaussian_blur)

        images_list, acts_list = [], []
        ignore_scale_scheduler = False
        last_scheduling_epoch = None
        self.learn_steps_taken = 0
        self.template.train()
        pbar = tqdm(range(num_max_epochs), bar_format=_BAR)
        epoch = -1  # num_max_epochs=0: no epoch runs, the call returns the snapshot of the untrained template like the reference
        for epoch in pbar:
            if save_results_every_n_epochs and epoch % save_results_every_n_epochs == 0:
                imgs, a = snapshot()
                images_list.append(imgs), acts_list.append(a)

            if fused:
                epoch_grids = torch.stack(list(dm.train_dataloader())).to(device, non_blocking=True)  # [steps, N, 1]: one H2D per epoch
                for i in range(epoch_grids.shape[0]):
                    stepper.step(epoch_grids[i])
                self.learn_steps_taken += int(epoch_grids.shape[0])
                acts = stepper.last_acts()
            else:
                for input_grid in dm.train_dataloader():
                    input_grid = input_grid.to(device)
                    _, img_post, _acts, _ = self.forward(input_grid, self.template, self.img_transforms, template_neuron_model,
                                                         return_template=True, gb=gaussian_blur)
                    acts = _acts / template_mei_act
                    sim = grid_reg.masked_reg_term(img_post, mask_dev, self.pixel_min, self.pixel_max)
                    loss = -acts.mean() - reg_scale * sim
                    optimizer.zero_grad()
                    loss.backward()
                    optimizer.step()
                    self.learn_steps_taken += 1

            if not ignore_scale_scheduler and epoch > min_epochs:
                # the scheduler monitors the LAST TRAINING STEP's mean activation, not the evaluation below (:212-215)
                new_scale = reg_scheduler.step(metric_to_monitor=acts.mean(), value_to_reduce=reg_scale)
                if fused and new_scale != reg_scale:
                    stepper.set_reg_scale(new_scale)
                reg_scale = new_scale
                with torch.no_grad():
                    if fused:
                        acts = stepper.evaluate(grid)
                    else:
                        _, _, _acts, _ = self.forward(grid, self.template, self.img_transforms, template_neuron_model, return_template=True,
                                                      gb=gaussian_blur)
                        acts = _acts / template_mei_act
                if check_activity_requirements(acts, requirements):
                    ignore_scale_scheduler = True
                    last_scheduling_epoch = epoch

            a_mean, a_min, a_std = torch.stack([acts.mean(), acts.min(), acts.std()]).tolist()  # one D2H for the three scalars
            desc = act_desc = f"Activation mean = {a_mean:.2f} (min = {a_min:.2f} std = {a_std:.2f})"
            if ignore_scale_scheduler and epoch >= (last_scheduling_epoch + additional_epochs):
                pbar.set_description("✅ Invariance learning is finished!" + " " + act_desc)
                pbar.close()
                break
            if verbose:
                desc += f" | Contrastive reg scale = {reg_scale:.3f} | Epochs w/o improving = {reg_scheduler.num_epochs_no_improvement} (patience = {reg_scheduler.patience})"
            pbar.set_description(desc, refresh=False)
        self.learn_epochs_run = epoch + 1

        self.template.eval()
        imgs, a = snapshot()
        if save_results_every_n_epochs is not None:
            images_list.append(imgs), acts_list.append(a)
            if torch.is_tensor(imgs):
                return torch.stack(images_list), torch.stack(acts_list)
            return np.stack(images_list), np.stack(acts_list)
        return imgs, a

    # ------------------------------------------------------------------------------------------------------------------
    # match
    # ------------------------------------------------------------------------------------------------------------------
    def match(self, target_neuron_idxs, grid_points_per_dim=20, steps_per_epoch=1, lr=1e-3, num_epochs=1000, patience=15,
              ignore_diff_smaller_than=1e-3, initialize_at_rf_positions=True, rotate_angle_and_scale=True, find_best_translation=False,
              only_affine_coordinate_transformation=True, stochastic_coordinate_transformation=False, allow_scale_coordinate_transformation=True,
              allow_shear_coordinate_transformation=True, uniform_scale_coordinate_transformation=False,
              init_noise_scale_coordinate_transformation=0.1, coordinate_transform_clamp_boundaries=True, verbose=False,
              save_results_every_n_epochs=None, _total_targets=None, _stop_reduce=None, _return_device=False):
        """`_total_targets` / `_stop_reduce` are used by laminr_b200.dist when the targets are sharded over GPUs: the loss keeps
        the global 1/D scale (:363) and the stop rule sees the global mean activation (:368)."""
        num_targets = len(target_neuron_idxs)
        self.target_neuron_idxs = target_neuron_idxs
        device = next(self.template.parameters()).device
        rf = {k: np.array([v["center_pos"]["x"], v["center_pos"]["y"]]).astype(np.float32) for k, v in self.meis_dict.items()}
        template_rf_mask = self.meis_dict[self.template_neuron_idx]["mask"]
        others_rf_mask = np.array([self.meis_dict[i]["mask"] for i in target_neuron_idxs])
        others_rf_location = np.array([rf[i] for i in target_neuron_idxs])
        others_mei_act = torch.from_numpy(np.array([self.meis_dict[i]["activation"] for i in target_neuron_idxs])).to(device)
        loc_in_model = target_neuron_idxs
        loc_in_list = np.arange(num_targets)

        self.template.initialize_coordinate_transform(
            num_neurons=num_targets, only_affine=only_affine_coordinate_transformation, stochastic=stochastic_coordinate_transformation,
            init_noise_scale=init_noise_scale_coordinate_transformation, allow_scale=allow_scale_coordinate_transformation,
            allow_shear=allow_shear_coordinate_transformation, uniform_scale=uniform_scale_coordinate_transformation,
            clamp_boundaries=coordinate_transform_clamp_boundaries)
        if initialize_at_rf_positions:
            self.template.register_coords_shifts(torch.from_numpy(rf[self.template_neuron_idx]).to(device), torch.from_numpy(others_rf_location).to(device))

        grid_dataloader = JitteringGridDatamodule(num_invariances=1, grid_points_per_dim=grid_points_per_dim, steps_per_epoch=steps_per_epoch)
        fused = self._fused_ok()
        stepper = None
        import os, sys, time
        timing, marks = os.environ.get("LMR_TIMING") == "1", []  # debug: per-phase wall clock on stderr (adds synchronisations)

        def mark(name):
            if timing:
                torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))

        mark("start")
        if fused:
            stepper = MatchStepper(self.template, self.img_transforms, self.response_predicting_model, loc_in_model, others_mei_act, grid_points_per_dim,
                                   lr=lr, total_targets=_total_targets, precision=self.precision, use_graph=self.use_cuda_graph)
            self._match_stepper = stepper

        mark("stepper construction")
        template = self.initialize_with_best_angle_and_translation(
            self.template, self.response_predicting_model, self.img_transforms, grid_dataloader, loc_in_list, loc_in_model,
            template_rf_mask=template_rf_mask, others_rf_mask=others_rf_mask, rotate_angle_and_scale=rotate_angle_and_scale,
            find_best_translation=find_best_translation, _stepper=stepper)

        mark("init sweep")
        if not fused:
            optimizer = torch.optim.Adam(template.coordinate_transform.parameters(), lr=lr)
        improvement_checker = ImprovementChecker(patience=patience, ignore_diff_smaller_than=ignore_diff_smaller_than)
        grid = grid_dataloader.grid.to(device)
        d_total = _total_targets if _total_targets is not None else num_targets

        # the final image block is large (0.8 GB for 1023 targets at 100x100): its host array is allocated NOW and made resident by a background
        # thread while the GPU works, so that the copy at the end runs at PCIe speed instead of at first-touch page-fault speed
        host_buf = None
        if fused and not _return_device and save_results_every_n_epochs is None and stepper.z.numel() * 4 >= (64 << 20):
            host_buf = ops.HostBuffer(stepper.z.numel(), np.float32)

        def snapshot(final=False):
            if fused:
                act = stepper.evaluate(grid)
                if _return_device:  # laminr_b200.dist collects the blocks of all ranks on the host itself
                    return stepper.z.clone(), act.clone()
                out = host_buf.wait() if (final and host_buf is not None) else None
                return ops.to_host_numpy(stepper.z, out=out), act.cpu().numpy()
            return self.get_matched_images_and_activations(grid, template, loc_in_list, loc_in_model)

        images_list, acts_list = [], []
        self.match_steps_taken = 0
        pbar = tqdm(range(num_epochs), bar_format=_BAR)
        epoch = -1  # num_epochs=0: the call returns the snapshot of the initialised transform like the reference
        epochs_run = 0

        def judge(a_sum, a_min, a_max):
            """The stop rule of :368-375 on one epoch's statistics -> (keep going?, progress-bar text)."""
            a_mean = a_sum / d_total
            going = improvement_checker.is_increasing(a_mean)
            return going, f"Activation mean = {a_mean:.2f} (min = {a_min:.2f} max = {a_max:.2f})"

        if fused:
            # Fused path: the statistics of epoch e are read while epoch e + 1 already runs -- the host's round trip (copy of 12 bytes, in the
            # sharded case a 3-float collective before it, the stop rule, the next launch) never leaves the GPU idle.  Epoch e + 1 is therefore
            # SPECULATIVE: its first step saves the state it starts from, and if the stop rule turns out to have fired at epoch e, that state is
            # restored, the CPU generator is rewound to before e + 1's jitter draw and a snapshot taken for e + 1 is dropped -- parameters,
            # results, epoch / step counts and generator state are those of the reference's sequential loop.
            pending = None  # (statistics handle of the previous epoch, its slot)
            stopped = False
            for epoch in pbar:
                took_snapshot = False
                if save_results_every_n_epochs and epoch % save_results_every_n_epochs == 0:
                    imgs, a = snapshot()
                    images_list.append(imgs), acts_list.append(a)
                    took_snapshot = True
                if not template.training:  # (the reference calls .train() every epoch, :350: a walk over the module tree; only the first one changes anything)
                    template.train()
                rng_before = torch.get_rng_state()
                epoch_grids = list(grid_dataloader.train_dataloader())  # same generator, same order as the reference (:351); CPU rows: the stepper
                #                                                        stages them in pinned memory (an asynchronous copy, see MatchStepper.step)
                slot = epoch & 1
                for i, input_grid in enumerate(epoch_grids):
                    stepper.step(input_grid, slot=slot, save=(i == 0))
                handle = _stop_reduce.begin(stepper.epoch2[slot]) if _stop_reduce is not None else stepper.queue_stats(slot)
                if pending is not None:
                    stats = _stop_reduce.end(pending[0]) if _stop_reduce is not None else stepper.wait_stats(pending[1])
                    going, act_desc = judge(*stats)
                    if not going:  # the PREVIOUS epoch was the last one: undo this one
                        stepper.restore_state()
                        torch.set_rng_state(rng_before)
                        if took_snapshot:
                            images_list.pop(), acts_list.pop()
                        if _stop_reduce is not None:
                            _stop_reduce.end(handle)  # (every rank leaves its last collective completed)
                        stopped = True
                        pbar.set_description("✅ Invariance matching is finished!" + " " + act_desc)
                        pbar.close()
                        break
                    desc = act_desc
                    if verbose:
                        desc += f" | Epochs w/o improving = {improvement_checker.has_not_improved_counter} (patience: {improvement_checker.patience})"
                    pbar.set_description(desc, refresh=False)
                self.match_steps_taken += len(epoch_grids)
                epochs_run = epoch + 1
                pending = (handle, slot)
            if not stopped and pending is not None:  # the last epoch's statistics: nothing left to undo, but the checker and the bar see them
                stats = _stop_reduce.end(pending[0]) if _stop_reduce is not None else stepper.wait_stats(pending[1])
                going, act_desc = judge(*stats)
                if not going:
                    pbar.set_description("✅ Invariance matching is finished!" + " " + act_desc)
                pbar.close()
        else:
            for epoch in pbar:
                if save_results_every_n_epochs and epoch % save_results_every_n_epochs == 0:
                    imgs, a = snapshot()
                    images_list.append(imgs), acts_list.append(a)
                template.train()
                for input_grid in grid_dataloader.train_dataloader():
                    input_grid = input_grid.to(device)
                    _, _, _acts, _ = self.forward(input_grid, template, self.img_transforms, self.response_predicting_model, return_template=False,
                                                  other_neurons_loc_in_list=loc_in_list, other_neurons_loc_in_model=loc_in_model)
                    relevant = _acts[range(num_targets), :, range(num_targets)]
                    acts = relevant.mean(dim=1) / others_mei_act
                    loss = -acts.sum() / d_total
                    optimizer.zero_grad()
                    loss.backward()
                    optimizer.step()
                    self.match_steps_taken += 1
                a_sum, a_min, a_max = torch.stack([acts.sum(), acts.min(), acts.max()]).tolist()
                if _stop_reduce is not None:
                    a_sum, a_min, a_max = _stop_reduce(a_sum, a_min, a_max)
                epochs_run = epoch + 1
                going, act_desc = judge(a_sum, a_min, a_max)
                if not going:
                    pbar.set_description("✅ Invariance matching is finished!" + " " + act_desc)
                    pbar.close()
                    break
                desc = act_desc
                if verbose:
                    desc += f" | Epochs w/o improving = {improvement_checker.has_not_improved_counter} (patience: {improvement_checker.patience})"
                pbar.set_description(desc, refresh=False)
        self.match_epochs_run = epochs_run
        mark(f"{epochs_run} epochs")

        template.eval()
        imgs, a = snapshot(final=True)
        mark("snapshot")
        if timing:
            print("[match] " + ", ".join(f"{b[0]} {1e3 * (b[1] - a_[1]):.1f} ms" for a_, b in zip(marks, marks[1:])), file=sys.__stderr__)
        if save_results_every_n_epochs is not None:
            images_list.append(imgs), acts_list.append(a)
            if torch.is_tensor(imgs):
                return torch.stack(images_list), torch.stack(acts_list)
            return np.stack(images_list), np.stack(acts_list)
        return imgs, a

    # ------------------------------------------------------------------------------------------------------------------
    # result getters (:394-422, :452-495)
    # ------------------------------------------------------------------------------------------------------------------
    def get_template_images_and_activations(self, grid, template_neuron_model, gaussian_blur):
        with torch.no_grad():
            _, img_post, _acts, _ = self.forward(grid, self.template, self.img_transforms, template_neuron_model, return_template=True, gb=gaussian_blur)
        return img_post.cpu().data.numpy(), _acts.squeeze().cpu().data.numpy()

    def get_matched_images_and_activations(self, input_grid, template, other_neurons_loc_in_list, other_neurons_loc_in_model):
        with torch.no_grad():
            _, img_post, _acts, _ = self.forward(input_grid, template, self.img_transforms, self.response_predicting_model, return_template=False,
                                                 other_neurons_loc_in_list=other_neurons_loc_in_list, other_neurons_loc_in_model=other_neurons_loc_in_model)
            n = len(self.target_neuron_idxs)
            relevant = _acts[range(n), :, range(n)]
        return img_post.cpu().data.numpy(), relevant.cpu().data.numpy()

    def _grid_for(self, n_images, device):
        return JitteringGridDatamodule(num_invariances=1, grid_points_per_dim=n_images, steps_per_epoch=1).grid.to(device)

    def get_images_from_learned_template(self, n_images=20):
        device = next(self.template.parameters()).device
        with torch.no_grad():
            return self.get_images(self._grid_for(n_images, device), self.template, self.img_transforms, return_template=True).cpu().data.numpy()

    def get_images_from_matched_template(self, target_neuron_idx, n_images=20):
        if not isinstance(target_neuron_idx, int):
            raise ValueError("target_neuron_idx should be an integer, referring to a single target neuron.")
        if target_neuron_idx not in self.target_neuron_idxs:
            raise ValueError(f"target_neuron_idx should be one of the target neuron indices: {self.target_neuron_idxs}")
        pos = int(np.where(target_neuron_idx == np.array(self.target_neuron_idxs))[0][0])
        device = next(self.template.parameters()).device
        channels, height, width = self.meis_dict[self.template_neuron_idx]["mei"].shape
        with torch.no_grad():
            images = self.get_images(self._grid_for(n_images, device), self.template, self.img_transforms, return_template=False)
            return images.reshape(len(self.target_neuron_idxs), n_images, channels, height, width)[pos].cpu().data.numpy()

    def save_learned_template_as_gif(self, *args, **kwargs):
        raise NotImplementedError("laminr_b200: GIF export is visualisation only (needs matplotlib / lipstick) and is out of scope")

    save_matched_template_as_gif = save_learned_template_as_gif

    # ------------------------------------------------------------------------------------------------------------------
    # generic differentiable pipeline (:497-571)
    # ------------------------------------------------------------------------------------------------------------------
    def get_images(self, grid, image_generator, img_transf, return_template=False, centralize_template=False, return_pre=False):
        _img_pre = image_generator(aux=grid, return_template=return_template, centralize_template=centralize_template)
        img_post = img_transf(_img_pre.flatten(0, -4))
        return (_img_pre, img_post) if return_pre else img_post

    def forward(self, grid, image_generator, img_transf, encoding_model, gb=None, return_template=False, other_neurons_loc_in_list=None,
                other_neurons_loc_in_model=None, centralize_template=False):
        """grid -> (img_pre, img_post, acts, img_post) with the reference's shapes."""
        _img_pre, img_post = self.get_images(grid, image_generator, img_transf, return_template=return_template,
                                             centralize_template=centralize_template, return_pre=True)
        if return_template:
            *N, C, H, W = _img_pre.shape
        else:
            D, *N, C, H, W = _img_pre.shape
        if gb is not None:
            img_post.register_hook(gb)
        single = getattr(encoding_model, "forward_neurons", None)
        if not return_template and single is not None:
            acts_sel = single(img_post, list(other_neurons_loc_in_model))  # only the requested neurons (same values as acts[:, idxs])
        else:
            acts_sel = None
            acts = encoding_model(img_post)
        img_pre = _img_pre.flatten(0, -4)
        if return_template:
            return img_pre.reshape(*N, C, H, W), img_post.reshape(*N, C, H, W), acts.reshape(*N, -1), img_post.reshape(*N, C, H, W)
        sel = acts_sel if acts_sel is not None else acts[:, other_neurons_loc_in_model]
        Dsel = len(other_neurons_loc_in_model)
        return (img_pre.reshape(D, *N, C, H, W)[other_neurons_loc_in_list], img_post.reshape(D, *N, C, H, W)[other_neurons_loc_in_list],
                sel.reshape(D, *N, Dsel)[other_neurons_loc_in_list], img_post.reshape(D, *N, C, H, W)[other_neurons_loc_in_list])

    # ------------------------------------------------------------------------------------------------------------------
    # match initialisation sweep (:573-650)
    # ------------------------------------------------------------------------------------------------------------------
    def initialize_with_best_angle_and_translation(self, template, encoding_model, img_transforms, grid_dataloader, other_neurons_loc_in_list,
                                                   other_neurons_loc_in_model, template_rf_mask=None, others_rf_mask=None,
                                                   rotate_angle_and_scale=True, find_best_translation=True, _stepper=None):
        device = next(template.parameters()).device
        aff = template.coordinate_transform.transforms.Affine
        n = len(other_neurons_loc_in_list)
        inv_fraction = None
        if rotate_angle_and_scale:
            if others_rf_mask is None:
                raise ValueError("others_rf_mask must be provided if rotate_angle_and_scale is True")
            fraction = torch.from_numpy(others_rf_mask.sum(axis=(1, 2)) / template_rf_mask.sum()).to(device)
            inv_fraction = (torch.ones_like(aff.scalings.data) * 1 / fraction.reshape(-1, 1)).to(aff.scalings.dtype)
        if not (find_best_translation or rotate_angle_and_scale):
            return template
        angles = np.linspace(0, 2 * np.pi, int(360 / 6 + 1))[:-1].astype(np.float32)
        translations = np.linspace(-0.1, 0.1, 11).astype(np.float32) if find_best_translation else np.array([0.0], np.float32)
        desc = ("Finding best angle and translation for initalization" if find_best_translation and rotate_angle_and_scale else
                "Finding best translation for initalization" if find_best_translation else "Finding best angle for initalization")
        table = np.zeros((len(angles), len(translations), len(translations), n))
        # fused path: the candidates' mean activations stay on the device and come back in ONE copy after the sweep (the reference reads them
        # back candidate by candidate, :641: 60 -- or 7,260 -- host synchronisations that leave the GPU idle between candidates)
        dev_table = torch.empty(table.shape, dtype=torch.float32, device=device) if _stepper is not None else None
        with torch.no_grad():
            input_grid = grid_dataloader.grid.to(device)
            if rotate_angle_and_scale:
                aff.scalings.data.copy_(inv_fraction)
            if _stepper is not None:
                # all candidates x all targets as virtual targets of a few batched launches; candidate order == the loop's (i, j, k)
                ii, jj, kk = np.meshgrid(np.arange(len(angles)), np.arange(len(translations)), np.arange(len(translations)), indexing="ij")
                cand_t = (translations[jj.ravel()], translations[kk.ravel()]) if find_best_translation else (None, None)
                dev_table.view(-1, n).copy_(_stepper.sweep(input_grid, angles[ii.ravel()], *cand_t))
            for i, angle in enumerate(tqdm(angles, desc=desc, bar_format=_BAR) if _stepper is None else ()):  # generic autograd path: the reference's loop
                aff.angles.data.fill_(float(angle))
                for j, tx in enumerate(translations):
                    for k, ty in enumerate(translations):
                        if find_best_translation:
                            aff.translation.data[:, 0, 0] = float(tx)
                            aff.translation.data[:, 0, 1] = float(ty)
                        _, _, _acts, _ = self.forward(input_grid, template, img_transforms, encoding_model, return_template=False,
                                                      other_neurons_loc_in_list=other_neurons_loc_in_list,
                                                      other_neurons_loc_in_model=other_neurons_loc_in_model)
                        table[i, j, k, :] = _acts[range(n), :, range(n)].mean(dim=1).cpu().numpy()
        if dev_table is not None:
            table[...] = dev_table.cpu().numpy()
        best = np.argmax(table.reshape(-1, n), axis=0)
        bi, bj, bk = np.unravel_index(best, table.shape[:3])
        aff.angles.data.copy_(torch.from_numpy(angles[bi]).to(device))
        aff.translation.data[:, 0, 0] = torch.from_numpy(translations[bj]).to(device)
        aff.translation.data[:, 0, 1] = torch.from_numpy(translations[bk]).to(device)
        if rotate_angle_and_scale:
            aff.scalings.data.copy_(inv_fraction)
        return template
