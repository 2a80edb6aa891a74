"""UCOD-DPL_dinov1: experiment-specific settings on top of ../__base__/shared_defaults.py."""

cfg = {'_BASE_': ['../__base__/shared_defaults.py'],
 'dataset_cfg': {'feature_extractor_cfg': {'backbone': 'facebook/dino-vitb8', 'type': 'dinov1'},
                 'trainloader_cfg': {'batch_size': 16},
                 'trainset_cfg': {'bkg_th': 0.3, 'image_size': (296, 296), 'require_label': False},
                 'valset_cfg': {'image_size': (296, 296)}},
 'exp_name': 'UCOD-DPL_dinov1',
 'model_cfg': {'ema_weight': 0.99},
 'train_cfg': {'lr0': 0.0006, 'max_epoch': 25, 'step_lr_size': 25},
 'val_cfg': {'look_twice_th': 0.05, 'val_interval': 5, 'val_start': 5}}
