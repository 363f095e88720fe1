"""TEST-SIDE stand-in: the one helper transformed_lm.py imports from human_body_prior (not installable here)."""
import os


def get_support_data_dir(current_fname=__file__):
    # the reference resolves <repo>/support_data relative to the calling file (src/moshpp/transformed_lm.py)
    d = os.path.dirname(os.path.abspath(current_fname))
    while d != os.path.dirname(d):
        cand = os.path.join(d, 'support_data')
        if os.path.isdir(cand):
            return cand
        d = os.path.dirname(d)
    raise FileNotFoundError('support_data')
