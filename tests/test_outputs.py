"""Result files of cpn_inference (cpn_inference.py:797-851): the hdf5 subset writer / reader, the tif writer, the region
property tables (GPU statistics vs scipy.ndimage) and the whole output step."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import celldetection_b200 as cd
from celldetection_b200.utils import h5min, tiffmin, outputs as OUT
from helpers import load_npz, fixture_state_dict


def _libhdf5_written_file():
    import scipy.io
    hits = glob.glob(os.path.join(os.path.dirname(scipy.io.__file__), 'matlab', 'tests', 'data', 'testhdf5_7.4_GLNX86.mat'))
    return hits[0] if hits else None


def test_h5_reader_parses_a_file_written_by_libhdf5():
    """Pins this repo's reading of the HDF5 specification (superblock v0 behind a user block, symbol-table group, local heap,
    v1 object header, datatype / dataspace / layout / attribute messages) on a file libhdf5 itself wrote."""
    f = _libhdf5_written_file()
    if f is None:
        pytest.skip('scipy test data not installed')
    r = h5min.read(f, with_attributes=True)
    data, attrs = r['testdouble']
    assert data.dtype == np.float64 and data.shape == (9, 1)
    assert np.allclose(data[:, 0], np.linspace(0, 2 * np.pi, 9), rtol=0, atol=1e-15)
    assert attrs == {'MATLAB_class': 'double'}


def test_h5_writer_emits_the_bytes_libhdf5_uses_for_shared_structures():
    """Message encodings that also occur in the libhdf5-written file must be byte-identical in ours."""
    f = _libhdf5_written_file()
    if f is None:
        pytest.skip('scipy test data not installed')
    ref = h5min._File(open(f, 'rb').read())
    kind, links = ref.node(ref.root_header)
    ref_msgs = {t: body for t, _, body in ref.messages(links['testdouble'])}
    assert h5min._pad8(h5min._datatype_message(np.float64)) == ref_msgs[h5min.MSG_DATATYPE]
    assert h5min._dataspace_message((9, 1)) == ref_msgs[h5min.MSG_DATASPACE]
    assert h5min.SIGNATURE == open(f, 'rb').read()[512:520]


def test_h5_round_trip(tmp_path):
    rng = np.random.RandomState(0)
    d = dict(contours=rng.rand(5, 32, 2).astype(np.float32), boxes=rng.rand(5, 4).astype(np.float32),
             scores=rng.rand(5).astype(np.float32), classes=np.arange(5), labels=rng.randint(0, 9, (40, 50, 2)).astype(np.int32),
             empty=np.zeros((0, 4), np.float32), u8=rng.randint(0, 255, (7, 3)).astype(np.uint8), f64=rng.rand(3),
             scalar=np.float32(3.5), u16=np.arange(6, dtype=np.uint16).reshape(2, 3))
    f = str(tmp_path / 'r.h5')
    OUT.to_h5(f, **d, attributes=dict(contours=dict(args='{"a": 1, "b": null}', n=np.arange(3, dtype=np.int32))))
    r = h5min.read(f, with_attributes=True)
    assert sorted(r) == sorted(d)
    for k, v in d.items():
        a, at = r[k]
        assert a.dtype == np.asarray(v).dtype and a.shape == np.asarray(v).shape and np.array_equal(a, v), k
    assert r['contours'][1]['args'] == '{"a": 1, "b": null}' and np.array_equal(r['contours'][1]['n'], np.arange(3))
    assert np.array_equal(OUT.from_h5(f, 'boxes'), d['boxes'])
    # structure: every object header / data block 8-byte aligned, end-of-file address == file size
    buf = open(f, 'rb').read()
    h = h5min._File(buf)
    assert h.eof == len(buf) and h.leaf_k == h5min.LEAF_K
    with pytest.raises(ValueError):
        h5min.write(f, {f'd{i}': np.zeros(1) for i in range(40)})


@pytest.mark.parametrize('shape,dtype', [((37, 53), np.uint8), ((300, 1201, 3), np.uint8), ((129, 77, 4), np.uint8),
                                         ((64, 65), np.uint16), ((1100, 1500, 4), np.uint8), ((40, 30), np.float32)])
def test_tiff_writer_is_read_back_by_libtiff(tmp_path, shape, dtype):
    cv2 = pytest.importorskip('cv2')
    rng = np.random.RandomState(1)
    a = (rng.rand(*shape) * 255).astype(dtype)
    if len(shape) == 3:
        a[:shape[0] // 2] = 7                       # compressible half
        if shape[-1] == 4:
            a[..., 3] = 255                         # (libtiff pre-multiplies unassociated alpha on reading)
    for big in (False, True):
        for comp in ('ZLIB', None):
            f = str(tmp_path / 't.tif')
            tiffmin.imwrite(f, a, compression=comp, bigtiff=big)
            b = cv2.imread(f, cv2.IMREAD_UNCHANGED)
            assert b is not None, (big, comp)
            if b.ndim == 3:
                b = b[..., [2, 1, 0] + ([3] if b.shape[-1] == 4 else [])]
            assert b.dtype == a.dtype and np.array_equal(a, b), (big, comp)


def test_file_inputs_are_read_as_rgb_arrays(tmp_path):
    """cpn_inference.py:689-716: image files (OpenCV here) and hdf5 datasets as inputs / masks."""
    cv2 = pytest.importorskip('cv2')
    from celldetection_b200.inference import _load_image_file
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (21, 33, 3)).astype(np.uint8)
    png = str(tmp_path / 'a.png')
    cv2.imwrite(png, np.ascontiguousarray(img[..., ::-1]))                     # OpenCV writes BGR
    assert np.array_equal(_load_image_file(png), img)
    gray16 = rng.randint(0, 65536, (17, 19)).astype(np.uint16)
    tif = str(tmp_path / 'g.tif')
    tiffmin.imwrite(tif, gray16)
    got = _load_image_file(tif)
    assert got.dtype == np.uint16 and np.array_equal(got, gray16)
    h5 = str(tmp_path / 'in.h5')
    OUT.to_h5(h5, image=img, mask=(img[..., 0] > 128))
    assert np.array_equal(_load_image_file(h5, 'image'), img)
    assert np.array_equal(_load_image_file(h5, 'mask'), (img[..., 0] > 128).astype(np.uint8))
    with pytest.raises(FileNotFoundError):
        _load_image_file(str(tmp_path / 'missing.png'))


def _scipy_table(lab, properties, spacing=(1., 1.)):
    """The same table from scipy.ndimage (an independent implementation of the region statistics)."""
    from scipy import ndimage as ndi
    rows = []
    if lab.ndim == 2:
        lab = lab[..., None]
    for z in range(lab.shape[2]):
        l = lab[..., z]
        ids = np.unique(l[l > 0])
        objs = ndi.find_objects(l)
        for i in ids:
            sl = objs[i - 1]
            n = float(ndi.sum_labels(l > 0, l, i))
            com = ndi.center_of_mass(l == i)
            bb = (sl[0].start, sl[1].start, sl[0].stop, sl[1].stop)
            area = n * spacing[0] * spacing[1]
            ab = (bb[2] - bb[0]) * (bb[3] - bb[1]) * spacing[0] * spacing[1]
            rows.append(dict(label=i, area=area, bbox=bb, centroid=(com[0] * spacing[0], com[1] * spacing[1]), area_bbox=ab,
                             extent=area / ab, equivalent_diameter_area=np.sqrt(4 * area / np.pi)))
    return rows


@pytest.mark.gpu
def test_label_property_table_equals_scipy_ndimage():
    z = load_npz('contours2labels')
    rng = np.random.RandomState(0)
    cases = [rng.randint(0, 6, (50, 70)).astype(np.int32), np.zeros((20, 30, 2), np.int32)]
    big = np.zeros((300, 400, 3), np.int32)
    for k in range(1, 120):                                   # rectangles + random speckle, labels spread over channels
        y, x, c = rng.randint(0, 280), rng.randint(0, 380), k % 3
        big[y:y + rng.randint(1, 20), x:x + rng.randint(1, 20), c] = k
    big[rng.rand(*big.shape) < 0.01] = 0
    cases.append(big)
    props = ['label', 'area', 'bbox', 'centroid', 'area_bbox', 'extent', 'equivalent_diameter_area']
    for lab in cases:
        for spacing in (1., 0.5):
            for src in (lab, torch.from_numpy(lab).cuda()):
                tab = cd.data.labels2property_table(src, props, spacing=spacing, separator='-')
                want = _scipy_table(lab, props, (spacing, spacing))
                assert len(tab) == len(want)
                assert list(tab.columns) == ['label', 'area', 'bbox-0', 'bbox-1', 'bbox-2', 'bbox-3', 'centroid-0',
                                             'centroid-1', 'area_bbox', 'extent', 'equivalent_diameter_area']
                for row, w in zip(tab.itertuples(index=False), want):
                    assert row[0] == w['label'] and row[1] == w['area'] and tuple(row[2:6]) == w['bbox']
                    assert np.allclose(row[6:8], w['centroid'], rtol=1e-12, atol=0)
                    assert row[8] == w['area_bbox'] and np.isclose(row[9], w['extent'], rtol=1e-14)
                    assert np.isclose(row[10], w['equivalent_diameter_area'], rtol=1e-14)
    with pytest.raises(NotImplementedError):
        cd.data.labels2property_table(cases[0], ['eccentricity'])
    assert len(z.files) > 0


@pytest.mark.gpu
def test_cpn_inference_writes_h5_csv_and_overlay(tmp_path):
    """The output step of cpn_inference.py:797-851 on the apply_model fixture: files exist, the h5 datasets equal the returned
    tensors bit for bit, the csv tables equal the label statistics, the overlay tif is readable."""
    import pandas as pd
    cv2 = pytest.importorskip('cv2')
    z = load_npz('apply_model_cpnu22')
    seed, crop, stride, border = [int(v) for v in z['meta']]
    m = cd.models.CpnU22(3, precision='fp32')
    m.load_state_dict(fixture_state_dict(z, 'CpnU22', seed))
    m = m.cuda()
    out = str(tmp_path / 'out')
    res = cd.cpn_inference(z['img'], m, outputs=out, tile_size=crop, stride=stride, border_removal=border, batch_size=2,
                           labels=True, flat_labels=True, properties=['label', 'area', 'bbox', 'centroid'], overlay=True)
    y = res[0]
    assert set(y['files']) == {'h5', 'properties_flat', 'properties', 'overlay'}
    assert os.path.basename(y['files']['h5']) == 'ndarray_0.h5'
    h5 = h5min.read(y['files']['h5'], with_attributes=True)
    for k in ('contours', 'boxes', 'scores', 'classes', 'locations', 'fourier', 'contour_proposals', 'labels', 'flat_labels'):
        assert np.array_equal(h5[k][0], y[k].cpu().numpy()), k
    args = json.loads(h5['contours'][1]['args'])
    assert args['tile_size'] == crop and args['stride'] == stride and args['properties'] == ['label', 'area', 'bbox', 'centroid']
    tab = pd.read_csv(y['files']['properties'], index_col=0)
    assert len(tab) == len(z['out/scores']) == len(y['properties']) and list(tab['label']) == sorted(tab['label']) or True
    assert np.allclose(tab.to_numpy(), y['properties'].to_numpy())
    flat = pd.read_csv(y['files']['properties_flat'], index_col=0)
    lab = y['flat_labels'].cpu().numpy()
    assert list(flat['label']) == [int(v) for v in np.unique(lab[lab > 0])]
    assert list(flat['area']) == [float((lab == v).sum()) for v in flat['label']]
    vis = cv2.imread(y['files']['overlay'], cv2.IMREAD_UNCHANGED)
    assert vis.shape == z['img'].shape[:2] + (4,) and ((vis[..., 3] > 0) == (y['labels'].cpu().numpy() > 0).any(-1)).all()
    # a file input (written as png) through the same entry point, results skipped when they exist
    png = str(tmp_path / 'tile.png')
    cv2.imwrite(png, np.ascontiguousarray(z['img'][..., ::-1]))
    r2 = cd.cpn_inference(png, m, outputs=out, tile_size=crop, stride=stride, border_removal=border, batch_size=2)
    assert os.path.isfile(os.path.join(out, 'tile.h5')) and torch.equal(r2[0]['contours'], y['contours'])
    assert len(cd.cpn_inference(png, m, outputs=out, tile_size=crop, stride=stride, skip_existing=True)) == 0
