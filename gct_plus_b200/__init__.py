"""gct_plus_b200 -- B200-native (sm_100a) implementation of the GCT-Plus Transformer-VAE hot path.

Mirrors the reference's package layout for the path it replaces:
    gct_plus_b200.Model.build_model           get_model / load_state / get_sampler / model_dict
    gct_plus_b200.Model.forward_propagation1  forward_propagation[model_type]
    gct_plus_b200.Train.trainer1              loss_function / KLAnnealer / run_epoch / train_model / FusedTrainer
    gct_plus_b200.Inference.sampling_tool     sampling_tool_dict[model_type]
All arithmetic runs in libgct_b200.so (include/gct_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
