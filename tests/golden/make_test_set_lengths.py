"""Extract the (mixture length, reference length) pairs of the reference's 3 000 full-length TSS test utterances
(datasets/tss/test_set.pkl: a pickled LibrimixSpe index - paths, lengths, reference choices; no audio) into
tests/golden/test_set_lengths.txt.  Index data only; run once in the build container where /root/reference exists:

    python tests/golden/make_test_set_lengths.py

The reference's test loop reads mixture idx at full length (segment=None) and, as enrolment, the file named in the
'reference' column, which is a source of ANOTHER test mixture of the 'min' Libri2Mix variant, i.e. as long as that
mixture (src/datasets/librimix_spe.py:40-62)."""
import os
import pickle
import sys
import types

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'test_set_lengths.txt')

sys.modules.setdefault('soundfile', types.ModuleType('soundfile'))      # imported by the dataset module, unused here
sys.path.insert(0, REF)
ds = pickle.load(open(os.path.join(REF, 'datasets', 'tss', 'test_set.pkl'), 'rb'))
df = ds.df
length = dict(zip(df['mixture_ID'], df['length']))
rows, missing = [], 0
for _, r in df.iterrows():
    rid = os.path.splitext(os.path.basename(r['reference']))[0]
    if rid not in length:
        missing += 1
    rows.append((int(r['length']), int(length.get(rid, r['length']))))
with open(OUT, 'w') as f:
    f.write('# mixture_samples reference_samples (8 kHz) of datasets/tss/test_set.pkl, in dataset order\n')
    for t, tr in rows:
        f.write(f'{t} {tr}\n')
ts = [t for t, _ in rows]
print(f'{len(rows)} utterances, {missing} references outside the index; mix min/mean/max {min(ts)}/{sum(ts) / len(ts):.0f}/{max(ts)}, '
      f'total {sum(ts) / 8000:.0f} audio-s')
