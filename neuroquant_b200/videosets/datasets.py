"""Frame dataset with the reference's interface (videosets/datasets.py): a sorted directory of PNGs,
read as float in [0,1], centre-cropped to (crop_h, crop_w); samples are {'img', 'idx', 'norm_idx'}.
Host-side I/O only -- the calibration loop keeps the decoded frames resident in HBM after the first
epoch (quantization/calib_model.py).  Additive: `as_uint8=True` hands the frames out as the PNGs store them (uint8);
the consumer evaluates value / 255 on the device (bit-identical, IEEE division), which quarters the host -> device
bytes and the HBM footprint of the resident clip."""
import os

from torch.utils.data import Dataset


class VideoDataSet(Dataset):
    def __init__(self, cfg, args, as_uint8: bool = False):
        self.as_uint8 = bool(as_uint8)
        self.video = [os.path.join(args.data_path, x) for x in sorted(os.listdir(args.data_path))]
        self.crop_h, self.crop_w = cfg["crop_h"], cfg["crop_w"]
        first = self.img_transform(self.img_load(0))
        if first.shape[0] != 3:
            raise ValueError(f"{self.video[0]}: expected 3-channel RGB frames, got {first.shape[0]} channels")
        self.final_size = first.size(-2) * first.size(-1)
        self.diff = cfg["diff_enc"]
        if self.diff:
            raise NotImplementedError("diff_enc datasets feed the FP training script only")

    def img_load(self, idx):
        from torchvision.io import read_image
        img = read_image(self.video[idx])
        return img if self.as_uint8 else img / 255.0

    def img_transform(self, img):
        from torchvision.transforms.functional import center_crop
        return center_crop(img, (self.crop_h, self.crop_w))

    def __len__(self):
        return len(self.video)

    def __getitem__(self, idx):
        return {"img": self.img_transform(self.img_load(idx)), "idx": idx, "norm_idx": float(idx) / len(self.video)}
