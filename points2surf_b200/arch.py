"""Parameter inventory of PointsToSurfModel (names / shapes / registration order), derived from the
module structure of source/points_to_surf_model.py:12-36,72-99,134-167,237-294."""


def layer_specs(use_point_stn, shared_transformation, net=1024, output_dim=2):
    """-> list of (dotted name, kind in {'conv','fc','bn'}, cout, cin) in state_dict order."""

    def stn(prefix, dim, quat):
        out = 4 if quat else dim * dim
        return [(prefix + 'conv1', 'conv', 64, dim), (prefix + 'conv2', 'conv', 128, 64),
                (prefix + 'conv3', 'conv', net, 128),
                (prefix + 'fc1', 'fc', net // 2, net), (prefix + 'fc2', 'fc', net // 4, net // 2),
                (prefix + 'fc3', 'fc', out, net // 4),
                (prefix + 'bn1', 'bn', 64, 0), (prefix + 'bn2', 'bn', 128, 0), (prefix + 'bn3', 'bn', net, 0),
                (prefix + 'bn4', 'bn', net // 2, 0), (prefix + 'bn5', 'bn', net // 4, 0)]

    def feat(prefix, point_stn):
        s = []
        if point_stn:
            s += stn(prefix + 'stn1.', 3, True)
        s += stn(prefix + 'stn2.', 64, False)
        s += [(prefix + 'conv0a', 'conv', 64, 3), (prefix + 'conv0b', 'conv', 64, 64),
              (prefix + 'bn0a', 'bn', 64, 0), (prefix + 'bn0b', 'bn', 64, 0),
              (prefix + 'conv1', 'conv', 64, 64), (prefix + 'conv2', 'conv', 128, 64),
              (prefix + 'conv3', 'conv', net, 128),
              (prefix + 'bn1', 'bn', 64, 0), (prefix + 'bn2', 'bn', 128, 0), (prefix + 'bn3', 'bn', net, 0)]
        return s

    specs = []
    if use_point_stn and shared_transformation:
        specs += stn('point_stn.', 3, True)
    specs += feat('feat_local.', False)
    specs += feat('feat_global.', bool(use_point_stn and not shared_transformation))
    specs += [('fc1_local', 'fc', net // 2, net), ('fc1_global', 'fc', net // 2, net),
              ('bn1_local', 'bn', net // 2, 0), ('bn1_global', 'bn', net // 2, 0),
              ('fc2', 'fc', net // 4, net), ('fc3', 'fc', net // 8, net // 4), ('fc4', 'fc', output_dim, net // 8),
              ('bn2', 'bn', net // 4, 0), ('bn3', 'bn', net // 8, 0)]
    return specs
