from .synth import synth_state_dict, calibrate_heads_
from .io import load_model, fetch_model, save_fetchable_model, model2dict, dict2model
from .outputs import to_h5, from_h5
