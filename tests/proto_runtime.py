"""types.proto (reference, /types.proto:1-46) as runtime-built python-protobuf
classes — there is no protoc in this image.  Test-side only: the product uses
the native codec (csrc/lfr_wire.cc)."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=_F.LABEL_OPTIONAL, type_name=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name
    return f


def build():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "types.proto"
    fd.syntax = "proto3"
    mf = fd.message_type.add()
    mf.name = "MatchingFile"
    ip = mf.nested_type.add()
    ip.name = "ImagePair"
    _field(ip, "image_name1", 1, _F.TYPE_STRING)
    _field(ip, "fact1", 2, _F.TYPE_FLOAT)
    _field(ip, "image_name2", 3, _F.TYPE_STRING)
    _field(ip, "fact2", 4, _F.TYPE_FLOAT)
    m = ip.nested_type.add()
    m.name = "Match"
    _field(m, "feature_idx1", 1, _F.TYPE_UINT32)
    _field(m, "feature_idx2", 2, _F.TYPE_UINT32)
    _field(m, "similarity", 3, _F.TYPE_FLOAT)
    d = m.nested_type.add()
    d.name = "Displacement"
    _field(d, "di", 1, _F.TYPE_FLOAT)
    _field(d, "dj", 2, _F.TYPE_FLOAT)
    _field(m, "disp1", 4, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".MatchingFile.ImagePair.Match.Displacement")
    _field(m, "disp2", 5, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".MatchingFile.ImagePair.Match.Displacement")
    _field(ip, "matches", 5, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".MatchingFile.ImagePair.Match")
    _field(mf, "image_pairs", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".MatchingFile.ImagePair")
    sf = fd.message_type.add()
    sf.name = "SolutionFile"
    im = sf.nested_type.add()
    im.name = "Image"
    _field(im, "image_name", 1, _F.TYPE_STRING)
    _field(im, "fact", 2, _F.TYPE_FLOAT)
    sd = im.nested_type.add()
    sd.name = "Displacement"
    _field(sd, "feature_idx", 1, _F.TYPE_UINT32)
    _field(sd, "di", 2, _F.TYPE_FLOAT)
    _field(sd, "dj", 3, _F.TYPE_FLOAT)
    _field(im, "displacements", 3, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".SolutionFile.Image.Displacement")
    _field(sf, "images", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".SolutionFile.Image")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return (message_factory.GetMessageClass(pool.FindMessageTypeByName("MatchingFile")),
            message_factory.GetMessageClass(pool.FindMessageTypeByName("SolutionFile")))


MatchingFile, SolutionFile = build()


def matchset_to_proto(ms, lo=0, hi=None):
    """The writer loop of compute_match_graph.py:163-187."""
    hi = ms.n_pairs if hi is None else hi
    mf = MatchingFile()
    for p in range(lo, hi):
        ip = mf.image_pairs.add()
        ip.image_name1 = ms.image_names[int(ms.pair_img1[p])]
        ip.fact1 = float(ms.pair_fact1[p])
        ip.image_name2 = ms.image_names[int(ms.pair_img2[p])]
        ip.fact2 = float(ms.pair_fact2[p])
        for m in range(int(ms.pair_ptr[p]), int(ms.pair_ptr[p + 1])):
            mm = ip.matches.add()
            mm.feature_idx1 = int(ms.feat1[m])
            mm.feature_idx2 = int(ms.feat2[m])
            mm.similarity = float(ms.sim[m])
            for g in range(9):
                d = mm.disp1.add()
                d.di = float(ms.disp1[m, 2 * g])
                d.dj = float(ms.disp1[m, 2 * g + 1])
                d = mm.disp2.add()
                d.di = float(ms.disp2[m, 2 * g])
                d.dj = float(ms.disp2[m, 2 * g + 1])
    return mf
