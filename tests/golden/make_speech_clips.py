"""Extract the four real-speech clips the reference embeds in example.ipynb (cell 15: mixture, target, the att
checkpoint's estimate, reference utterance; 24 000 samples, 8 kHz, int16, each peak-normalised by IPython.display.Audio)
into ``notebook_clips.npz``, and produce real-speech reference fixtures from the LIVE reference on them.

Run in the build container only (needs /root/reference):

    python tests/golden/make_speech_clips.py

SURVEY.md section 4 names these clips as the only real-speech pin the reference carries.  The checkpoint that produced
the embedded estimate is missing, so the estimate is stored as a sanity signal only; the fixtures are the reference's own
modules (seeded default init; and with the LSTM weights scaled x3, i.e. gates driven into saturation, as a stand-in for
trained weights) run on the real mixture / reference utterance.
"""
import base64
import io
import json
import os
import re
import sys
import wave

import numpy as np
import torch

sys.path.insert(0, '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import BASE, fingerprint  # noqa: E402

NAMES = ('mix', 'target', 'estimate', 'reference')


def extract():
    nb = json.load(open('/root/reference/example.ipynb'))
    clips = []
    for o in nb['cells'][15]['outputs']:
        if o['output_type'] != 'display_data':
            continue
        m = re.search(r'src="data:audio/wav;base64,([^"]+)"', ''.join(o['data']['text/html']))
        w = wave.open(io.BytesIO(base64.b64decode(m.group(1))))
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 8000, 24000)
        clips.append(np.frombuffer(w.readframes(24000), dtype='<i2').copy())
    assert len(clips) == 4
    np.savez_compressed(os.path.join(HERE, 'notebook_clips.npz'), **dict(zip(NAMES, clips)))
    return dict(zip(NAMES, clips))


def scale_lstm(model, s):
    with torch.no_grad():
        for n, p in model.named_parameters():
            if '.rnn.weight_' in n:
                p.mul_(s)


def speech_case(name, clips, fusion, lstm_wscale, wseed=0):
    from src.models.dprnn_spe import DPRNNSpeTasNet
    kwargs = dict(BASE, n_repeats=6, fusion_type=fusion)
    torch.manual_seed(wseed)
    model = DPRNNSpeTasNet(**kwargs).eval()
    scale_lstm(model, lstm_wscale)
    # soundfile's float32 normalisation of int16 PCM (what LibrimixSpe feeds the model): x / 32768
    mix = torch.from_numpy(clips['mix'].astype(np.float32) / 32768.0)[None]
    ref = torch.from_numpy(clips['reference'].astype(np.float32) / 32768.0)[None]
    with torch.no_grad():
        est, logits = model(mix, ref, torch.tensor(24000.))
    meta = dict(cls='src.models.dprnn_spe.DPRNNSpeTasNet', kwargs=kwargs, B=1, T=24000, Tr=24000, training=False,
                wseed=wseed, iseed=-1, lstm_wscale=lstm_wscale, weight_fingerprint=fingerprint(model.state_dict()),
                torch=torch.__version__, data='example.ipynb cell 15 (real speech)')
    np.savez_compressed(os.path.join(HERE, name + '.npz'), meta=json.dumps(meta), mix=mix.numpy(), ref=ref.numpy(),
                        est=est.numpy(), logits=logits.numpy())
    print(name, 'fp', meta['weight_fingerprint'], 'est peak', float(est.abs().max()))


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count())
    clips = extract()
    speech_case('speech_att_r6', clips, 'att', 1.0)          # the notebook's own model configuration
    speech_case('speech_cat_r6_wx3', clips, 'cat', 3.0)      # cfg 2's model, saturated gates
