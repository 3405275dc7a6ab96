"""Hot-path subset of the reference's configuration bag.

Attribute names and defaults follow utils/config.py:14-408 of the reference so that either
this class or the reference's own ``Config`` object can be handed to Decoder / NeuralPoints /
Mapper (they only read attributes).  ``load`` understands the YAML sections that touch the
hot path (utils/config.py:410-910); everything else in a run file is ignored.
"""
from __future__ import annotations

import torch


class Config:
    def __init__(self) -> None:
        # runtime
        self.device: str = "cuda"
        self.dtype = torch.float32
        self.tran_dtype = torch.float64
        self.seed: int = 42
        self.silence: bool = True
        self.wandb_vis_on: bool = False
        # ranges (process section)
        self.min_range: float = 2.5
        self.max_range: float = 60.0
        # neural points (utils/config.py:110-131)
        self.voxel_size_m: float = 0.3
        self.weighted_first: bool = True
        self.layer_norm_on: bool = False
        self.num_nei_cells: int = 2
        self.query_nn_k: int = 6
        self.use_mid_ts: bool = False
        self.search_alpha: float = 0.2
        self.idw_index: int = 2
        self.buffer_size: int = int(5e7)
        self.local_voxel_size_m: float = 0.2    # LocalPointCloudMap (utils/config.py:111,124,148)
        self.local_buffer_size: int = int(5e6)
        self.local_map_size: float = 100.0
        self.feature_dim: int = 8
        self.feature_std: float = 0.0
        self.color_on: bool = False
        self.color_channel: int = 0
        self.semantic_on: bool = False
        self.local_map_travel_dist_ratio: float = 5.0
        self.diff_ts_local: float = 400.0
        self.prune_map_on: bool = False
        self.prune_freq_frame: int = 100
        self.max_prune_certainty: float = 3.0
        self.from_sample_points: bool = True
        self.from_all_samples: bool = False
        self.map_surface_ratio: float = 0.5
        self.pool_filter_freq: int = 1
        # sampler
        self.surface_sample_range_m: float = 0.25
        self.surface_sample_n: int = 3
        self.free_sample_begin_ratio: float = 0.3
        self.free_sample_end_dist_m: float = 1.0
        self.free_front_n: int = 2
        self.free_behind_n: int = 1
        # decoder (utils/config.py:167-189)
        self.mlp_bias_on: bool = True
        self.mlp_leaky_relu: bool = False
        self.geo_mlp_level: int = 1
        self.geo_mlp_hidden_dim: int = 64
        self.freeze_after_frame: int = 40
        self.use_gaussian_pe: bool = False
        self.pos_encoding_band: int = 0
        self.pos_input_dim: int = 3
        # loss (utils/config.py:191-222)
        self.main_loss_type: str = "bce"
        self.sigma_sigmoid_m: float = 0.1
        self.logistic_gaussian_ratio: float = 0.55
        self.proj_correction_on: bool = False
        self.loss_weight_on: bool = False
        self.dist_weight_on: bool = True
        self.dist_weight_scale: float = 0.8
        self.numerical_grad: bool = True
        self.gradient_decimation: int = 10
        self.num_grad_step_ratio: float = 0.2
        self.ekional_loss_on: bool = True
        self.ekional_add_to: str = "all"
        self.weight_e: float = 0.5
        self.consistency_loss_on: bool = False
        # optimiser (utils/config.py:224-241)
        self.iters: int = 12
        self.init_iter_ratio: int = 40
        self.opt_adam: bool = True
        self.bs: int = 16384
        self.lr: float = 0.01
        self.weight_decay: float = 0.0
        self.adam_eps: float = 1e-15
        self.adaptive_iters: bool = False
        self.new_sample_ratio_less: float = 0.02
        self.new_sample_ratio_more: float = 0.15
        self.new_sample_ratio_restart: float = 0.3
        # replay pool
        self.bs_new_sample: int = 2048
        self.new_certainty_thre: float = 1.0
        self.pool_capacity: int = int(1e7)
        # tracking / pgo switches the mapper looks at
        self.track_on: bool = True          # utils/config.py:254; load() keeps it only if the run file has a tracker section
        self.pgo_on: bool = False
        self.use_pin_mapper: bool = False   # utils/config.py:18: CLID-SLAM's region-specific sampler is the default
        self.consistency_count: int = 1000  # utils/config.py:218-219 (consistency loss; not on the fused path)
        self.consistency_range: float = 0.05
        self.lr_pose: float = 1e-4          # utils/config.py:235 (bundle adjustment; no caller)
        self.dynamic_certainty_thre: float = 0.5
        self.dynamic_sdf_ratio_thre: float = 1.5
        self.dynamic_min_grad_norm_thre: float = 0.25
        self._derive()

    def _derive(self) -> None:  # utils/config.py:902-910
        self.infer_bs = self.bs * 64
        self.window_radius = max(self.max_range, 6.0)
        self.local_map_radius = self.max_range + 2.0

    def load(self, path: str) -> None:
        import yaml

        with open(path, "r") as fh:
            args = yaml.safe_load(fh)
        setting = args.get("setting", {}) or {}
        self.use_pin_mapper = bool(setting.get("use_pin_mapper", False))      # utils/config.py:415
        self.track_on = bool(args.get("tracker", False))                      # utils/config.py:676: on only if indicated
        self.pgo_on = bool(args.get("pgo", False)) if self.track_on else False  # utils/config.py:742-743
        proc = args.get("process", {})
        self.min_range = proc.get("min_range_m", self.min_range)
        self.max_range = proc.get("max_range_m", self.max_range)
        vox_down_m = proc.get("vox_down_m", self.max_range * 1e-3)
        smp = args.get("sampler", {})
        self.local_voxel_size_m = smp.get("local_voxel_size_m", vox_down_m)
        self.surface_sample_range_m = smp.get("surface_sample_range_m", vox_down_m * 3.0)
        self.surface_sample_n = smp.get("surface_sample_n", self.surface_sample_n)
        self.free_sample_begin_ratio = smp.get("free_sample_begin_ratio", self.free_sample_begin_ratio)
        self.free_sample_end_dist_m = smp.get("free_sample_end_dist_m", self.surface_sample_range_m * 4.0)
        self.free_front_n = smp.get("free_front_sample_n", self.free_front_n)
        self.free_behind_n = smp.get("free_behind_sample_n", self.free_behind_n)
        npt = args.get("neuralpoints", {})
        self.voxel_size_m = npt.get("voxel_size_m", vox_down_m * 5.0)
        self.weighted_first = npt.get("weighted_first", self.weighted_first)
        self.layer_norm_on = npt.get("layer_norm_on", self.layer_norm_on)
        self.num_nei_cells = npt.get("num_nei_cells", self.num_nei_cells)
        self.query_nn_k = npt.get("query_nn_k", self.query_nn_k)
        self.search_alpha = npt.get("search_alpha", self.search_alpha)
        self.feature_dim = npt.get("feature_dim", self.feature_dim)
        self.prune_map_on = npt.get("prune_map_on", self.prune_map_on)
        dec = args.get("decoder", {})
        self.mlp_leaky_relu = dec.get("mlp_leaky_relu", self.mlp_leaky_relu)
        self.geo_mlp_level = dec.get("mlp_level", self.geo_mlp_level)
        self.geo_mlp_hidden_dim = dec.get("mlp_hidden_dim", self.geo_mlp_hidden_dim)
        self.freeze_after_frame = dec.get("freeze_after_frame", self.freeze_after_frame)
        loss = args.get("loss", {})
        self.main_loss_type = loss.get("main_loss_type", self.main_loss_type)
        self.sigma_sigmoid_m = loss.get("sigma_sigmoid_m", vox_down_m)
        self.loss_weight_on = loss.get("loss_weight_on", self.loss_weight_on)
        self.dist_weight_scale = loss.get("dist_weight_scale", self.dist_weight_scale)
        self.ekional_loss_on = loss.get("ekional_loss_on", self.ekional_loss_on)
        self.weight_e = float(loss.get("weight_e", self.weight_e))
        self.numerical_grad = loss.get("numerical_grad_on", self.numerical_grad)
        if not self.numerical_grad:
            self.gradient_decimation = 1
        else:
            self.gradient_decimation = loss.get("grad_decimation", self.gradient_decimation)
            self.num_grad_step_ratio = loss.get("num_grad_step_ratio", self.num_grad_step_ratio)
        cont = args.get("continual", {})
        self.bs_new_sample = int(cont.get("batch_size_new_sample", self.bs_new_sample))
        self.pool_capacity = int(float(cont.get("pool_capacity", self.pool_capacity)))
        self.new_certainty_thre = float(cont.get("new_certainty_thre", self.new_certainty_thre))
        opt = args.get("optimizer", {})
        self.iters = opt.get("iters", self.iters)
        self.bs = opt.get("batch_size", self.bs)
        self.lr = float(opt.get("learning_rate", self.lr))
        self.weight_decay = float(opt.get("weight_decay", self.weight_decay))
        self.adaptive_iters = opt.get("adaptive_iters", self.adaptive_iters)
        self.lr_pose = float(opt.get("lr_pose_ba", self.lr_pose))
        self._derive()
        self.consistency_count = int(self.bs / 4)  # utils/config.py:904


def ncd128() -> Config:
    """Shapes of config/run_ncd128.yaml (the shipped Newer-College run file)."""
    cfg = Config()
    cfg.max_range = 60.0
    cfg.min_range = 1.0
    cfg.surface_sample_range_m = 0.25
    cfg.surface_sample_n = 4
    cfg.free_sample_begin_ratio = 0.5
    cfg.free_sample_end_dist_m = 1.2
    cfg.free_front_n = 2
    cfg.voxel_size_m = 0.4
    cfg.num_nei_cells = 2
    cfg.search_alpha = 0.5
    cfg.weighted_first = True
    cfg.sigma_sigmoid_m = 0.1
    cfg.loss_weight_on = True
    cfg.dist_weight_scale = 0.8
    cfg.bs_new_sample = 1000
    cfg.pool_capacity = int(1e7)
    cfg.iters = 10
    cfg.bs = 16384
    cfg.lr = 0.01
    cfg.adaptive_iters = True
    cfg._derive()
    return cfg
