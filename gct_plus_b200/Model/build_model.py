"""Drop-in for Model/build_model.py: model factory, checkpoint loader, sampler factory."""
import os
from collections import OrderedDict

import torch

from . import Cvaetf, Vaetf

model_dict = {
    'vaetf': Vaetf,
    'pvaetf': Cvaetf,
    'scavaetf': Cvaetf,
    'pscavaetf': Cvaetf,
}


def extract_params(args, src_vocab_len, trg_vocab_len):
    return {
        'src_vocab': src_vocab_len, 'trg_vocab': trg_vocab_len, 'N': args.N, 'd_model': args.d_model, 'dff': args.d_ff,
        'h': args.H, 'latent_dim': args.latent_dim, 'dropout': args.dropout, 'use_cond2dec': args.use_cond2dec,
        'use_cond2lat': args.use_cond2lat, 'nconds': len(args.property_list), 'get_attn': args.get_attn,
    }


def load_state(model, model_path, rank, allow_pickle=False):
    """Accepts {'model_state_dict': ...} checkpoints or a bare state_dict, with or without DDP's
    'module.' prefix (reference build_model.py:59-76).  Checkpoints written by save_checkpoint hold tensors and plain
    containers only, so they load with weights_only=True; `allow_pickle=True` (or GCT_B200_ALLOW_PICKLE=1) opts in to
    full unpickling for legacy files from a trusted source."""
    allow_pickle = allow_pickle or os.environ.get('GCT_B200_ALLOW_PICKLE') == '1'
    ckpt = torch.load(model_path, map_location=torch.device('cpu'), weights_only=not allow_pickle)
    model_state = ckpt['model_state_dict'] if isinstance(ckpt, dict) and 'model_state_dict' in ckpt else ckpt
    if list(model_state.keys())[0].split('.')[0] == 'module':
        model_state = OrderedDict((k[7:], v) for k, v in model_state.items())
    model.load_state_dict(model_state)
    return model


def get_model(args, src_vocab_len, trg_vocab_len, rank):
    model = model_dict[args.model_type](**extract_params(args, src_vocab_len, trg_vocab_len))
    if hasattr(args, 'compute_dtype'):
        model.set_compute_dtype(args.compute_dtype)
    if hasattr(args, 'pad_id'):
        model.pad_id = args.pad_id
    if hasattr(args, 'model_path'):
        model = load_state(model, args.model_path, rank)
    return model


def get_sampler(args, SRC, TRG, toklen_data, scaler, device):
    from ..Inference.sampling_tool import sampling_tool_dict
    args.model_path = os.path.join(args.model_folder, args.model_name)
    model = get_model(args, len(SRC.vocab), len(TRG.vocab), device)
    model = model.to(device)
    model.eval()
    print(f'#parameters: {sum(p.numel() for p in model.parameters())}')
    kwargs = {
        'top_k': args.top_k, 'latent_dim': args.latent_dim, 'max_strlen': args.max_strlen,
        'use_cond2dec': args.use_cond2dec, 'decode_algo': args.decode_algo, 'n_jobs': args.n_jobs,
        'toklen_data': toklen_data, 'cond_dim': len(args.property_list), 'scaler': scaler, 'device': device,
        'SRC': SRC, 'TRG': TRG,
    }
    return sampling_tool_dict[args.model_type](model, kwargs)
