"""CORAL_dinov2: experiment-specific settings on top of ../__base__/shared_defaults.py."""

cfg = {'_BASE_': ['../__base__/shared_defaults.py'],
 'dataset_cfg': {'feature_extractor_cfg': {'backbone': 'facebook/dinov2-base', 'type': 'dinov2'},
                 'trainloader_cfg': {'batch_size': 2},
                 'trainset_cfg': {'bkg_th': 0.6,
                                  'image_size': (518, 518),
                                  'look_twice': False,
                                  'look_twice_th': 0.15,
                                  'require_label': True,
                                  'require_m_patches': True,
                                  'use_cache': True},
                 'valset_cfg': {'image_size': (518, 518), 'require_m_patches': False, 'use_cache': True}},
 'enable_plabel_cache': True,
 'exp_name': 'UCOD-DPL_dinov2',
 'model_cfg': {'ema_weight': 0.7, 'threshold': 0.0015, 'window_length': 56, 'window_size': 3},
 'start_ema': 1,
 'train_cfg': {'lr0': 0.0001, 'max_epoch': 8, 'step_lr_size': 2},
 'val_cfg': {'look_twice_th': 0.15, 'val_interval': 4, 'val_start': 4}}
