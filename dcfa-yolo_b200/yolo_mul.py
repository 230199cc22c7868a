"""Drop-in for the reference's prediction facade (yolo_mul.py:16-260): same class name, defaults, constructor
keywords and `detect_image / get_FPS / get_map_txt` signatures, backed by the sm_100a library -- plus the batched
entry point the reference lacks:

    yolo = YOLO(model_path='best_epoch_weights.pth', phi='s')
    results = yolo.detect_images(list_of_rgb_pil_images, list_of_depth_pil_images)
    # results[i]: None or float32 (n_i, 6) rows (top, left, bottom, right, score, class) in ORIGINAL image pixels

Per batch: ONE upload of the images' own uint8 pixels (any sizes; a single-channel depth frame stays one plane), the
reference's letterbox (PIL BICUBIC resize + grey padding, utils/utils.py:24-37) bit-exactly on the device, forward (the stem
kernel folds preprocess_input, the HWC->CHW transpose and cvtColor's replication of the depth plane) + decode + NMS + the
un-letterbox to each image's own pixel frame (utils/utils_bbox.py:60-85) on the device, ONE download of the kept rows.
`device_letterbox=False` keeps the PIL letterbox on the host.  `model_path=None` keeps the constructor's random
initialisation (tests, benchmarks).  Heat-map rendering and ONNX export are out of scope.
"""
import colorsys
import os
import time

import numpy as np
import torch
from PIL import ImageDraw, ImageFont

from nets.yolo_mul import YoloBody
from utils.utils import cvtColor, get_classes, letterbox_batch, show_config
from utils.utils_bbox import DecodeBox

_HERE = os.path.dirname(os.path.abspath(__file__))


class YOLO(object):
    _defaults = {
        "model_path": 'model_data/best_epoch_weights.pth',
        "classes_path": 'model_data/voc_classes.txt',
        "input_shape": [640, 640],
        "phi": 'n',
        "confidence": 0.5,
        "nms_iou": 0.3,
        "letterbox_image": True,
        "cuda": True,
        "device_letterbox": True,     # extension: resize_image on the GPU (bit-exact with PIL); False = PIL on the host
    }

    @classmethod
    def get_defaults(cls, n):
        if n in cls._defaults:
            return cls._defaults[n]
        return "Unrecognized attribute name '" + n + "'"

    def __init__(self, **kwargs):
        self.__dict__.update(self._defaults)
        for name, value in kwargs.items():
            setattr(self, name, value)
            self._defaults[name] = value
        classes_path = self.classes_path
        if not os.path.isabs(classes_path) and not os.path.exists(classes_path):
            classes_path = os.path.join(_HERE, classes_path)      # the copy shipped next to this file
        self.class_names, self.num_classes = get_classes(classes_path)
        self.bbox_util = DecodeBox(self.num_classes, (self.input_shape[0], self.input_shape[1]))
        self.bbox_util.iou_mode = 'cuda'   # the reference with cuda=True calls torchvision's CUDA nms kernel
        hsv = [(i / self.num_classes, 1.0, 1.0) for i in range(self.num_classes)]
        self.colors = [tuple(int(c * 255) for c in colorsys.hsv_to_rgb(*t)) for t in hsv]
        self._pinned = {}
        self.generate()
        show_config(**self._defaults)

    def generate(self, onnx=False):
        if onnx:
            raise NotImplementedError("dcfa_b200: ONNX export is out of scope (the forward runs in a CUDA library)")
        if not self.cuda or not torch.cuda.is_available():
            raise RuntimeError("dcfa_b200 has no CPU path: YOLO needs cuda=True and an sm_100a device")
        self.net = YoloBody(self.input_shape, self.num_classes, self.phi)
        if self.model_path:
            self.net.load_state_dict(torch.load(self.model_path, map_location='cpu'))
            print('{} model, and classes loaded.'.format(self.model_path))
        else:
            print('no model_path: constructor-initialised weights, classes loaded.')
        self.net = self.net.eval().cuda()   # no nn.DataParallel: one process per GPU

    # ------------------------------------------------------------------------------------------ batched path
    def _upload(self, images, tag, allow_plane=False):
        """-> (uint8 CUDA [B,H,W,3] (or [B,H,W] for single-channel frames when allow_plane), original (h, w) shapes [B,2])."""
        b = len(images)
        h, w = int(self.input_shape[0]), int(self.input_shape[1])
        if not self.device_letterbox:   # PIL letterbox on the host into a reused pinned buffer, one async copy
            key = (tag, b, h, w)
            host = self._pinned.get(key)
            if host is None:
                host = self._pinned[key] = torch.empty(b, h, w, 3, dtype=torch.uint8).pin_memory()
            _, shapes = letterbox_batch(images, (h, w), self.letterbox_image, out=host.numpy())
            return host.cuda(non_blocking=True), shapes
        # the images' own pixels in ONE pinned staging buffer and one async copy, then resize_image on the device
        from dcfa_b200.engine import letterbox_u8
        arrs = [np.asarray(im) for im in images]
        plane = allow_plane and all(a.ndim == 2 for a in arrs)
        if not plane:
            arrs = [a if (a.ndim == 3 and a.shape[2] == 3) else np.asarray(cvtColor(im)) for a, im in zip(arrs, images)]
        arrs = [np.ascontiguousarray(a, dtype=np.uint8) for a in arrs]
        shapes = np.array([a.shape[:2] for a in arrs], dtype=np.int64)
        offs = np.concatenate([[0], np.cumsum([(a.size + 255) // 256 * 256 for a in arrs])])
        host = self._pinned.get(tag)
        if host is None or host.numel() < offs[-1]:
            host = self._pinned[tag] = torch.empty(int(offs[-1] * 1.5) + 256, dtype=torch.uint8).pin_memory()
        hv = host.numpy()
        for a, o in zip(arrs, offs):
            hv[o:o + a.size] = a.reshape(-1)
        dev = host[:int(offs[-1])].cuda(non_blocking=True)
        out = torch.empty((b, h, w) if plane else (b, h, w, 3), dtype=torch.uint8, device=dev.device)
        for i, (a, o) in enumerate(zip(arrs, offs)):
            src = dev[int(o):int(o) + a.size].view(a.shape)
            self._lb_ws = letterbox_u8(src, out[i], self.letterbox_image, getattr(self, '_lb_ws', None))
        return out, shapes

    def detect_images(self, images_rgb, images_nir):
        """Lists of PIL images (same length, same size per pair) -> list of None | (n_i, 6) float32 rows
        (top, left, bottom, right, score, class) in each image's own pixel coordinates."""
        if len(images_rgb) != len(images_nir) or len(images_rgb) == 0:
            raise ValueError("detect_images needs two non-empty lists of the same length")
        rgb, shapes = self._upload(images_rgb, 'rgb')
        nir, _ = self._upload(images_nir, 'nir', allow_plane=True)
        with torch.no_grad():
            outputs = self.bbox_util.decode_box(self.net(rgb, nir))
            return self.bbox_util.non_max_suppression(outputs, self.num_classes, self.input_shape, shapes,
                                                      self.letterbox_image, conf_thres=self.confidence,
                                                      nms_thres=self.nms_iou)

    # ------------------------------------------------------------------------------------------ reference API
    def detect_image(self, image_rgb, image_nir):
        """One pair in, the RGB image with the detections drawn on it out (what predict_mul.py shows or saves)."""
        image_rgb = cvtColor(image_rgb)
        result = self.detect_images([image_rgb], [image_nir])[0]
        if result is None:
            return image_rgb
        labels = result[:, 5].astype('int32')
        scores, boxes = result[:, 4], result[:, :4]
        size = int(np.floor(3e-2 * image_rgb.size[1] + 0.5))
        try:
            font = ImageFont.truetype(font='model_data/simhei.ttf', size=size)
        except OSError:
            font = ImageFont.load_default()
        thickness = int(max((image_rgb.size[0] + image_rgb.size[1]) // np.mean(self.input_shape), 1))
        draw = ImageDraw.Draw(image_rgb)
        for c, box, score in zip(labels, boxes, scores):
            top, left, bottom, right = box
            top, left = max(0, int(np.floor(top))), max(0, int(np.floor(left)))
            bottom = min(image_rgb.size[1], int(np.floor(bottom)))
            right = min(image_rgb.size[0], int(np.floor(right)))
            label = '{} {:.2f}'.format(self.class_names[int(c)], score)
            x0, y0, x1, y1 = draw.textbbox((0, 0), label, font=font)
            tw, th = x1 - x0, y1 - y0
            print(label.encode('utf-8'), top, left, bottom, right)
            origin = (left, top - th) if top - th >= 0 else (left, top + 1)
            for i in range(thickness):
                draw.rectangle([left + i, top + i, right - i, bottom - i], outline=self.colors[int(c)])
            draw.rectangle([origin, (origin[0] + tw, origin[1] + th)], fill=self.colors[int(c)])
            draw.text(origin, label, fill=(0, 0, 0), font=font)
        del draw
        return image_rgb

    def get_FPS(self, image_rgb, image_nir, test_interval):
        """Seconds per pair of the device path (forward + decode + NMS + result fetch) on an already uploaded pair."""
        rgb, shapes = self._upload([image_rgb], 'rgb')
        nir, _ = self._upload([image_nir], 'nir')

        def once():
            with torch.no_grad():
                outputs = self.bbox_util.decode_box(self.net(rgb, nir))
                return self.bbox_util.non_max_suppression(outputs, self.num_classes, self.input_shape, shapes,
                                                          self.letterbox_image, conf_thres=self.confidence,
                                                          nms_thres=self.nms_iou)
        once()
        torch.cuda.synchronize()
        t1 = time.time()
        for _ in range(test_interval):
            once()
        torch.cuda.synchronize()
        return (time.time() - t1) / test_interval

    def get_map_txt(self, image_id, image_rgb, image_nir, class_names, map_out_path):
        """detection-results/<image_id>.txt in the format get_map_mul.py reads: class score left top right bottom."""
        result = self.detect_images([image_rgb], [image_nir])[0]
        with open(os.path.join(map_out_path, "detection-results/" + image_id + ".txt"), "w", encoding='utf-8') as f:
            if result is None:
                return
            for row in result:
                top, left, bottom, right = row[:4]
                name = self.class_names[int(row[5])]
                if name not in class_names:
                    continue
                f.write("%s %s %s %s %s %s\n" % (name, str(row[4])[:6], str(int(left)), str(int(top)), str(int(right)),
                                                 str(int(bottom))))

    def detect_heatmap(self, image_rgb, image_nir, heatmap_save_path):
        raise NotImplementedError("dcfa_b200: heat-map rendering is out of scope")

    def convert_to_onnx(self, simplify, model_path):
        raise NotImplementedError("dcfa_b200: ONNX export is out of scope")
