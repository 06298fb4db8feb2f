"""Host-side logic that needs no GPU: batch sharding and the plane maps that replace the reference's channel
shuffle / sort copies (property tests)."""
import torch
from hypothesis import given, settings, strategies as st

from oracle import restate as R


@given(n=st.integers(0, 257), world=st.integers(1, 16))
@settings(max_examples=200, deadline=None)
def test_shard_range_partitions_the_batch(n, world):
    from rpst.dist import shard_range
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (lo, hi), (lo2, _) in zip(spans, spans[1:]):
        assert lo <= hi == lo2
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@given(n=st.integers(1, 4), groups_of=st.integers(1, 8), seed=st.integers(0, 10_000))
@settings(max_examples=100, deadline=None)
def test_plane_maps_reproduce_shuffle_and_sort(n, groups_of, seed):
    import rpst
    c = 4 * groups_of
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, 2, 3, generator=g)
    att = torch.rand(n, c, 1, 1, generator=g)
    shuf, srt = rpst.shuffle_map(n, c, 4), rpst.sort_map(att)
    take = lambda m: x.flatten(0, 1)[m.long()].view_as(x)
    assert torch.equal(take(shuf), R.channel_shuffle(x))                     # network/adain_rp.py:304-311
    assert torch.equal(take(srt), R.sort_by_weights(x, att))                 # network/adain_rp.py:230-249
    # sort applied after shuffle == one composed map; maps stay inside their own sample
    assert torch.equal(take(rpst.compose_maps(shuf, srt)), R.sort_by_weights(R.channel_shuffle(x), att))
    for m in (shuf, srt):
        assert m.dtype == torch.int32 and sorted(m.tolist()) == list(range(n * c))
        assert torch.equal(m.view(n, c) // c, torch.arange(n)[:, None].expand(n, c).to(torch.int32))
    assert rpst.compose_maps(None, srt) is srt and rpst.compose_maps(shuf, None) is shuf
