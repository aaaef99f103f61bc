"""Room driver (configs[4]): sharding over ranks renders every window exactly once, and a window rendered inside a
multi-window launch equals the same window rendered alone."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_room_shards_cover_all_windows_and_batches_are_independent(cuda_device):
    from spsg_b200 import room as R, synthetic as S
    dims = (64, 96, 128)
    room = R.synthetic_room_sdf(dims, cuda_device, seed=5)
    predict = R.synthetic_predictor(room)
    kw = dict(views_per_chunk=2, width=96, height=64, max_num_locs_per_sample=200000, keep_images=True)
    whole = R.render_room(predict, dims, cuda_device, chunks_per_launch=4, **kw)
    assert whole["windows"] == 3 * 4 and whole["rendered_windows"] > 0
    assert whole["label_hist"].sum() == whole["rays"]
    # two "ranks" run one after the other on this GPU: together they render what one rank renders
    parts = [R.render_room(predict, dims, cuda_device, chunks_per_launch=3, rank=r, world=2, **kw) for r in range(2)]
    assert parts[0]["rendered_windows"] + parts[1]["rendered_windows"] == whole["rendered_windows"]
    np.testing.assert_array_equal(parts[0]["label_hist"] + parts[1]["label_hist"], whole["label_hist"])
    got = {w: img for p in parts for (w, img) in p["images"]}
    for w, img in whole["images"]:
        assert torch.equal(got[w], img), "window %s differs between launch groupings" % (w,)
    # something was actually hit and labelled
    assert whole["label_hist"][:S.NUM_CLASSES].sum() > 0.1 * whole["rays"]
