"""A stand-in for the reference's ROS2 node that makes exactly the calls and attribute reads the
node makes on the mapper (rclpy is not installed, so scripts/3d_mapper_node.py itself cannot be
imported).  Every step cites the node line it replays.  The harness is mapper-agnostic: it is run
once on the unmodified reference (tests/golden/make_golden.py, build container) and once on the
drop-in shim (tests/test_fake_node.py, GPU box), and the two transcripts are compared.
"""
import importlib.util
import struct

import numpy as np


def load_mapper_class(path):
    """scripts/3d_mapper_node.py:37-42: the node execs the file `3d_mapper.py` next to itself."""
    spec = importlib.util.spec_from_file_location("mapper_3d", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.SonarTo3DMapper


def node_config(p):
    """scripts/3d_mapper_node.py:118-146: the 19-key dict, mount angles converted degrees -> radians."""
    return {
        'horizontal_fov': p['horizontal_fov'], 'vertical_aperture': p['vertical_aperture'],
        'max_range': p['max_range'], 'min_range': p['min_range'], 'intensity_threshold': p['intensity_threshold'],
        'sonar_position': [p['sonar_position.x'], p['sonar_position.y'], p['sonar_position.z']],
        'sonar_orientation': [np.radians(p['sonar_orientation.roll']), np.radians(p['sonar_orientation.pitch']),
                              np.radians(p['sonar_orientation.yaw'])],
        'voxel_resolution': p['voxel_resolution'], 'min_probability': p['min_probability'],
        'dynamic_expansion': p['dynamic_expansion'], 'z_filter_min': p['z_filter_min'],
        'z_filter_enabled': p['z_filter_enabled'], 'adaptive_update': p['adaptive_update'],
        'adaptive_threshold': p['adaptive_threshold'], 'adaptive_max_ratio': p['adaptive_max_ratio'],
        'log_odds_occupied': p['log_odds_occupied'], 'log_odds_free': p['log_odds_free'],
        'log_odds_min': p['log_odds_min'], 'log_odds_max': p['log_odds_max'],
    }


# the node's declared defaults (scripts/3d_mapper_node.py:53-107) with the shipped YAML's sensor values
NODE_PARAMS = {
    'horizontal_fov': 70.0, 'vertical_aperture': 20.0, 'max_range': 10.0, 'min_range': 1.0,
    'intensity_threshold': 120, 'sonar_position.x': 0.0, 'sonar_position.y': 0.0, 'sonar_position.z': -0.1,
    'sonar_orientation.roll': 0.0, 'sonar_orientation.pitch': 60.0, 'sonar_orientation.yaw': 0.0,
    'voxel_resolution': 0.15, 'min_probability': 0.7, 'dynamic_expansion': True, 'z_filter_min': -6.3,
    'z_filter_enabled': True, 'adaptive_update': True, 'adaptive_threshold': 0.5, 'adaptive_max_ratio': 0.3,
    'log_odds_occupied': 0.5, 'log_odds_free': -0.1, 'log_odds_min': -10.0, 'log_odds_max': 7.0,
}


class FakeNode:
    def __init__(self, mapper_cls, params, show_free_space):
        self.show_free_space = show_free_space
        self.mapper = mapper_cls(node_config(params))                       # :163
        self.frame_count = 0
        self.log = []          # what get_logger().info would print (without wall-clock fields)
        self.clouds = []       # PointCloud2 payloads (bytes)
        self.markers = []      # MarkerArray contents
        self.log.append(f"threshold={self.mapper.intensity_threshold}")     # :260 (banner / visualisation read it)

    def synchronized_callback(self, image, encoding, position, orientation):
        if encoding in ('mono8', '8UC1'):                                     # :305-306
            sonar_image = image
        elif encoding in ('mono16', '16UC1'):                                 # :307-310
            sonar_image = (image / 256).astype(np.uint8)
        else:
            self.log.append(f"Unsupported image encoding: {encoding}")       # :312
            return
        position = [float(position[0]), float(position[1]), float(position[2])]          # :319-323
        orientation = [float(orientation[0]), float(orientation[1]), float(orientation[2]), float(orientation[3])]
        stats = self.mapper.process_sonar_image(sonar_image, position, orientation)     # :333
        self.frame_count += 1
        if not stats.get('skipped', False) and self.frame_count % 10 == 0:               # :345
            assert stats["processing_time"] >= 0.0                                        # :356
            self.log.append(f'Frame {self.frame_count}: {stats["num_occupied"]} occupied, '
                            f'{stats["num_free"]} free, {stats["num_voxels"]} total voxels')   # :351-354

    def publish_pointcloud(self):
        result = self.mapper.get_point_cloud(include_free=self.show_free_space)        # :396
        if self.show_free_space:
            self.publish_marker_array(result)                                            # :400
        elif result['num_occupied'] > 0:                                                 # :403
            self.publish_pointcloud2(result['points'], result['probabilities'])         # :404

    def publish_pointcloud2(self, points, probabilities):
        data = []
        for i in range(len(points)):                                                     # :437-441
            data.append(struct.pack('ffff', points[i, 0], points[i, 1], points[i, 2], probabilities[i]))
        self.clouds.append({"width": len(points), "data": b''.join(data)})              # :424, :443

    def publish_marker_array(self, result):
        out = {}
        scale = self.mapper.voxel_resolution                                             # :466-468
        if len(result['occupied']) > 0:                                                  # :459
            pts = []
            for point, prob in result['occupied']:                                       # :474-476
                x, y, z = point
                pts.append((float(x), float(y), float(z)))
            out['occupied'] = {"scale": scale, "rgba": (1.0, 0.0, 0.0, 0.8), "points": pts}
        if self.show_free_space and len(result['free']) > 0:                             # :482
            pts = []
            for point, prob in result['free']:                                           # :497-499
                x, y, z = point
                pts.append((float(x), float(y), float(z)))
            out['free'] = {"scale": scale, "rgba": (0.0, 0.0, 1.0, 0.3), "points": pts}
        if len(result.get('unknown', [])) > 0:                                           # :505
            pts = []
            for point, prob in result['unknown']:                                        # :520-522
                x, y, z = point
                pts.append((float(x), float(y), float(z)))
            out['unknown'] = {"scale": scale, "rgba": (1.0, 1.0, 0.0, 0.5), "points": pts}
        self.markers.append(out)

    def shutdown(self):
        result = self.mapper.get_point_cloud()                                           # :543
        self.log.append(f'Final statistics: Total frames: {result["frame_count"]} Processed frames: '
                        f'{result["processed_count"]} Total voxels: {result["num_voxels"]} '
                        f'Occupied voxels: {result["num_occupied"]}')                    # :544-549


def run_session(mapper_cls, images, images16, pos, quat, show_free_space, publish_every=5):
    """A bag replay: every frame goes through the callback (frames listed in `images16` arrive as
    mono16), the 10 Hz publish timer fires every `publish_every` frames.  Returns the transcript with
    order-independent point sets (the reference's order is dict insertion order)."""
    node = FakeNode(mapper_cls, NODE_PARAMS, show_free_space)
    for f in range(len(images)):
        if f in images16:
            node.synchronized_callback(images16[f], 'mono16', pos[f], quat[f])
        else:
            node.synchronized_callback(images[f], 'mono8', pos[f], quat[f])
        if (f + 1) % publish_every == 0:
            node.publish_pointcloud()
    node.synchronized_callback(images[0], 'rgb8', pos[0], quat[0])                       # unsupported encoding (:311-313)
    node.publish_pointcloud()
    node.shutdown()
    clouds = []
    for c in node.clouds:
        a = np.frombuffer(c["data"], dtype='<f4').reshape(-1, 4)
        assert len(a) == c["width"]
        clouds.append(a[np.lexsort(a.T[::-1])])
    markers = []
    for m in node.markers:
        rec = {}
        for name, mk in m.items():
            p = np.asarray(mk["points"], dtype=np.float64).reshape(-1, 3)
            rec[name] = {"scale": mk["scale"], "rgba": list(mk["rgba"]), "points": p[np.lexsort(p.T[::-1])]}
        markers.append(rec)
    return {"log": node.log, "clouds": clouds, "markers": markers}
