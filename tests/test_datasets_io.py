"""CPU: host readers of the dataset formats (MetaImage, Duke-style .mat, npy/npz) round-trip synthetic files,
and the oracle's boundary rasteriser inverts boundary extraction on layered maps."""
import numpy as np
import pytest

from oracle import labelmap_oracle as lo
from retinal_oct_image_segmentation_via_deep_learning_b200 import datasets, synth


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32])
@pytest.mark.parametrize("compressed", [False, True])
def test_mhd_round_trip(tmp_path, dtype, compressed):
    rng = np.random.default_rng(3)
    vol = rng.integers(0, 9, size=(5, 12, 20)).astype(dtype)
    p = tmp_path / "vol.mhd"
    datasets.write_mhd(str(p), vol, spacing=(0.01, 0.002, 0.05), compressed=compressed)
    back, hdr = datasets.read_mhd(str(p))
    assert back.dtype == vol.dtype and np.array_equal(back, vol)
    assert hdr["DimSize"] == "20 12 5" and hdr["NDims"] == "3"


def test_mhd_local_data_and_big_endian(tmp_path):
    vol = np.arange(2 * 3 * 4, dtype=np.int16).reshape(2, 3, 4)
    p = tmp_path / "local.mhd"
    head = ("ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = True\n"
            "DimSize = 4 3 2\nElementType = MET_SHORT\nElementDataFile = LOCAL\n")
    p.write_bytes(head.encode() + vol.astype(">i2").tobytes())
    back, _ = datasets.read_mhd(str(p))
    assert np.array_equal(back, vol) and back.dtype == np.int16


def test_mhd_errors(tmp_path):
    p = tmp_path / "bad.mhd"
    p.write_text("ObjectType = Image\nNDims = 2\nDimSize = 4 4\nElementType = MET_UCHAR\nElementDataFile = missing.raw\n")
    with pytest.raises(FileNotFoundError):
        datasets.read_mhd(str(p))
    (tmp_path / "short.raw").write_bytes(b"\0" * 3)
    p.write_text("ObjectType = Image\nNDims = 2\nDimSize = 4 4\nElementType = MET_UCHAR\nElementDataFile = short.raw\n")
    with pytest.raises(ValueError):
        datasets.read_mhd(str(p))


def test_mat_layers_duke_layout(tmp_path):
    from scipy.io import savemat
    rng = np.random.default_rng(5)
    layers = rng.uniform(50, 400, size=(8, 64, 6))          # boundaries x columns x B-scans
    layers[:, :10, :] = np.nan                                # unannotated columns
    layers[:, :, 4] = np.nan                                  # an unannotated B-scan
    images = rng.integers(0, 255, size=(496, 64, 6)).astype(np.uint8)
    p = tmp_path / "Subject_01.mat"
    savemat(str(p), {"manualLayers1": layers, "images": images})
    got, imgs = datasets.read_mat_layers(str(p))
    assert got.shape == (6, 8, 64) and got.dtype == np.float32
    np.testing.assert_array_equal(got, np.transpose(layers, (2, 0, 1)).astype(np.float32))
    assert imgs.shape == (6, 496, 64)
    assert datasets.annotated_scans(got).tolist() == [0, 1, 2, 3, 5]
    with pytest.raises(KeyError):
        datasets.read_mat_layers(str(p), key="nope")


def test_read_labels_npy_npz(tmp_path):
    lab = np.random.default_rng(1).integers(0, 4, size=(3, 8, 8))
    np.save(tmp_path / "a.npy", lab)
    np.savez(tmp_path / "b.npz", seg=lab.astype(np.int32))
    for name in ("a.npy", "b.npz"):
        got = datasets.read_labels(str(tmp_path / name))
        assert got.dtype == np.uint8 and np.array_equal(got, lab)


def test_oracle_rasteriser_inverts_boundary_rows():
    yt, _ = synth.layered_pair(3, 60, 48, 5, seed=9)
    rows = np.stack([(yt < k).sum(axis=1) for k in range(1, 5)], axis=1)     # [n, K-1, W] = #{label < k} per column
    assert np.array_equal(lo.labels_from_boundaries(rows, 60), yt)
    shuffled = rows[:, ::-1, :].astype(np.float32) - 0.25                   # any order, fractional positions
    assert np.array_equal(lo.labels_from_boundaries(shuffled, 60), yt)


def _write_vol(path, bscans, seg, slo=(16, 8), hdr_size=None):
    """A Heidelberg .vol file with the layout read_heidelberg_vol documents (synthetic: no real file offline)."""
    import struct
    n, size_z, size_x = bscans.shape
    n_seg = seg.shape[1]
    off_seg = 256
    hdr_size = hdr_size or off_seg + n_seg * size_x * 4
    head = bytearray(2048)
    head[0:11] = b"HSF-OCT-103"
    struct.pack_into("<iii", head, 12, size_x, n, size_z)
    struct.pack_into("<ddd", head, 24, 0.0056, 0.12, 0.0039)
    struct.pack_into("<ii", head, 48, slo[0], slo[1])
    struct.pack_into("<i", head, 100, hdr_size)
    out = bytes(head) + bytes(slo[0] * slo[1])
    for i in range(n):
        bh = bytearray(hdr_size)
        bh[0:10] = b"HSF-BS-103"
        struct.pack_into("<i", bh, 12, hdr_size)
        struct.pack_into("<ii", bh, 48, n_seg, off_seg)
        s = np.where(np.isnan(seg[i]), np.float32(3.4028235e38), seg[i]).astype("<f4")
        bh[off_seg:off_seg + s.nbytes] = s.tobytes()
        img = np.where(np.isnan(bscans[i]), np.float32(3.4028235e38), bscans[i]).astype("<f4")
        out += bytes(bh) + img.tobytes()
    with open(path, "wb") as f:
        f.write(out)


def test_heidelberg_vol_round_trip(tmp_path):
    from retinal_oct_image_segmentation_via_deep_learning_b200 import datasets
    rng = np.random.default_rng(9)
    bscans = rng.random((3, 32, 48)).astype(np.float32)
    bscans[1, 0, :5] = np.nan                                     # samples outside the scan cone
    seg = np.sort(rng.uniform(2, 30, size=(3, 2, 48)), axis=1).astype(np.float32)
    seg[2, 1, 10:14] = np.nan                                     # boundary not found in these columns
    _write_vol(tmp_path / "x.vol", bscans, seg)
    got = datasets.read_heidelberg_vol(tmp_path / "x.vol")
    np.testing.assert_array_equal(got["bscans"], bscans)
    np.testing.assert_array_equal(got["boundaries"], seg)
    assert got["header"]["SizeX"] == 48 and got["header"]["NumBScans"] == 3 and got["scale"][0] == 0.0056
    with open(tmp_path / "bad.vol", "wb") as f:
        f.write(b"not a vol file" + bytes(3000))
    with pytest.raises(ValueError):
        datasets.read_heidelberg_vol(tmp_path / "bad.vol")
    with open(tmp_path / "x.vol", "rb") as f:
        blob = f.read()
    with open(tmp_path / "short.vol", "wb") as f:
        f.write(blob[:-100])
    with pytest.raises(ValueError):
        datasets.read_heidelberg_vol(tmp_path / "short.vol")


def test_hcms_delineation(tmp_path):
    from scipy.io import savemat
    from retinal_oct_image_segmentation_via_deep_learning_b200 import datasets
    rng = np.random.default_rng(10)
    bd = np.sort(rng.uniform(50, 400, size=(64, 5, 9)), axis=2)          # columns x B-scans x boundaries, as on disk
    savemat(tmp_path / "d.mat", {"bd_pts": bd, "control_pts": np.zeros((2, 2))})
    got = datasets.read_hcms_delineation(tmp_path / "d.mat")
    assert got.shape == (5, 9, 64) and got.dtype == np.float32
    np.testing.assert_allclose(got[3, 7], bd[:, 3, 7].astype(np.float32))
    savemat(tmp_path / "e.mat", {"other": bd})
    with pytest.raises(KeyError):
        datasets.read_hcms_delineation(tmp_path / "e.mat")
