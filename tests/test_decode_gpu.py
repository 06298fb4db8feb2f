"""Drop-in decode()/test() mirrors (rpst.decode) against the outputs of the reference's OWN methods
(tests/golden/decode.npz: network/adain_rp.py:251-269, 286-302, 538-553, 780-799 bound to the stub modules
of oracle/gen_golden.py).  This pins the integration layer — level order, blend vs concat, the decoder
state used as content, sort/shuffle as plane maps — not just the kernels."""
import types

import pytest
import torch

from oracle import restate as R
from oracle.gen_golden import decode_inputs, decode_stub

pytestmark = pytest.mark.gpu
TOL = 1e-5     # conv layers in fp32 on both sides; the transform itself sits at ~1e-6


def _stub(golden, kind):
    g = golden("decode")
    m = decode_stub(kind)
    m.load_state_dict({k[len(kind) + 1:]: v for k, v in g.items() if k.startswith(kind + ".") and
                       k[len(kind) + 1:] in m.state_dict()})
    cs, ss, atts = decode_inputs(kind)
    return g, m.cuda(), [c.cuda() for c in cs], [s.cuda() for s in ss], atts


def test_multiscale_decode_sort_and_shuffled_test(golden):
    from rpst import decode as D
    g, m, cs, ss, atts = _stub(golden, "multiscale")
    with torch.no_grad():
        assert R.rel_l2(D.decode_multiscale(m, cs, ss), g["multiscale.out"]) < TOL
        for enc, a in zip(m.rp_shared_encoder, atts):
            enc.attention_map = a.cuda()
        m._sort = True
        assert R.rel_l2(D.decode_multiscale(m, cs, ss), g["multiscale.out_sorted"]) < TOL
        m._shuffle = True
        feats = {"c": cs, "s": ss}
        m.encode_rp_intermediate = lambda x: feats[x]
        assert R.rel_l2(D.test_multiscale(m, "c", "s"), g["multiscale.test_out"]) < TOL
    assert m.training      # test() restores train mode like the reference (network/adain_rp.py:268)


def test_ldms_decode_uses_the_decoder_state_as_content(golden):
    from rpst import decode as D
    g, m, cs, ss, _ = _stub(golden, "ldms")
    with torch.no_grad():
        assert R.rel_l2(D.decode_ldms(m, cs, ss), g["ldms.out"]) < TOL


def test_ld_concat_decode(golden):
    from rpst import decode as D
    g, m, cs, ss, _ = _stub(golden, "ldcat")
    with torch.no_grad():
        assert R.rel_l2(D.decode_ld_concat(m, cs, ss), g["ldcat.out"]) < TOL


def test_multiscale_decode_with_masks_and_sort(golden):
    """use_mask=True through the decode mirror: segment AdaIN at every level with the running decoder state
    blended in the same call (`prev`), the sort permutation materialised first (the segment op takes dense
    tensors).  Expected values restated with the oracle's seg_adain_batch on the reference's loop
    (network/adain_rp.py:286-302, 313-319)."""
    from rpst import decode as D
    g, m, cs, ss, atts = _stub(golden, "multiscale")
    n, _, h, w = cs[0].shape
    cl = R.synth_labels(n, h, w, classes=3, block=4, seed=4400)
    sl = R.synth_labels(n, h, w, classes=3, block=4, seed=5400)
    for enc, a in zip(m.rp_shared_encoder, atts):
        enc.attention_map = a.cuda()
    m._sort = True
    with torch.no_grad():
        got = D.decode_multiscale(m, cs, ss, use_mask=True, c_mask_path=cl.cuda(), s_mask_path=sl.cuda())
        m64 = m.double().cpu()
        csd = [R.sort_by_weights(c.cpu().double(), a.double()) for c, a in zip(cs, atts)]
        ssd = [R.sort_by_weights(s.cpu().double(), a.double()) for s, a in zip(ss, atts)]
        st = m64.rp_decoder[0](R.seg_adain_batch(csd[-1], ssd[-1], cl, sl, dtype=torch.float64))
        for i, l in enumerate((1, 0)):
            st = m64.rp_decoder[i + 1](st + R.seg_adain_batch(csd[l], ssd[l], cl, sl, dtype=torch.float64))
    assert R.rel_l2(got, st) < TOL
