"""Drop-in for Model/forward_propagation1.py: builds the masks on the device and calls
model.forward with trg = batch['trg'][:, :-1]."""
from .modules import get_src_mask, get_trg_mask


def _no_conds(model, batch, pad_id, use_cond2dec):
    trg_in = batch['trg'][:, :-1]
    return model.forward(src=batch['src'], trg=trg_in, src_mask=get_src_mask(batch['src'], pad_id),
                         trg_mask=get_trg_mask(trg_in, pad_id, use_cond2dec))


def vaetf_forward_propagation(model, batch, pad_id, use_cond2dec):
    # the reference also prints len(outputs) here (forward_propagation1.py:14); omitted on purpose
    return _no_conds(model, batch, pad_id, use_cond2dec)


def scavaetf_forward_propagation(model, batch, pad_id, use_cond2dec):
    return _no_conds(model, batch, pad_id, use_cond2dec)


def pvaetf_forward_propagation(model, batch, pad_id, use_cond2dec):
    trg_in = batch['trg'][:, :-1]
    return model.forward(src=batch['src'], trg=trg_in,
                         src_mask=get_src_mask(batch['src'], pad_id, batch['econds']),
                         trg_mask=get_trg_mask(trg_in, pad_id, use_cond2dec, batch['dconds']),
                         econds=batch['econds'], dconds=batch['dconds'])


forward_propagation = {
    'vaetf': vaetf_forward_propagation,
    'scavaetf': scavaetf_forward_propagation,
    'pvaetf': pvaetf_forward_propagation,
    'pscavaetf': pvaetf_forward_propagation,
}
