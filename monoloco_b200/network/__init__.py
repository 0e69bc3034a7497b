from .net import Loco
from .process import load_calibration, factory_for_gt, preprocess_pifpaf, unnormalize_bi, extract_outputs, \
    extract_labels, extract_labels_aux
