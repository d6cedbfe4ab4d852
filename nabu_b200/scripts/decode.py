"""`python -m nabu_b200.scripts.decode --expdir <dir>` (reference: nabu/scripts/decode.py:13-60): decode the sections
recognizer.cfg names with the trained model of the experiment, results in <expdir>/decoded."""
import argparse

from . import load_model, read_cfg
from ..neuralnetworks.recognizer import Recognizer


def decode(expdir, testing=False, device='cuda'):
    database_cfg = read_cfg(expdir, 'database.conf', 'database.cfg')
    recognizer_cfg = read_cfg(expdir, 'recognizer.cfg')
    model = load_model(expdir)
    model.device = device
    recognizer = Recognizer(model=model, conf=recognizer_cfg, dataconf=database_cfg, expdir=expdir)
    if testing:
        return recognizer
    recognizer.recognize()
    return recognizer


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--expdir', default='expdir', help='the experiments directory that was used for training')
    decode(ap.parse_args().expdir, False)
