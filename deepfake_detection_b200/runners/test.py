"""`test_img(model_path, img_files)` with the reference's behaviour (dfd/runners/test.py:29-60): load
`model_half.pth.tar` into `efficientnet_deepfake_v4` (12 input channels = 4 RGB frames), fp16 eval forward, softmax,
column 0 = fake score.  The frame is replicated 4x on the channel axis exactly as the reference does (test.py:56)."""
import numpy as np
import torch

from ..models import create_deepfake_model_v4
from ..params import DeepFakeModel, img_mean, img_std, padding_image, resize


def test_img(model_path, img_files):
    from PIL import Image
    model = create_deepfake_model_v4("efficientnet_deepfake_v4", num_classes=2, in_chans=12, checkpoint_path=model_path,
                                     strict=False, dtype="fp16")
    model = DeepFakeModel(model)
    model.eval()
    scores_out = []
    for img_file in img_files:
        img = np.transpose(padding_image(resize(np.array(Image.open(img_file).convert("RGB"), np.uint8))), (2, 0, 1))
        img = torch.from_numpy(img).float().sub_(img_mean).div_(img_std).cuda().half()
        clip = torch.cat([img, img.clone(), img.clone(), img.clone()], dim=0).unsqueeze(0)
        with torch.no_grad():
            scores = model(clip)
        s = scores.float().cpu().numpy()[:, 0].tolist()
        print("{}'s fake score:{}".format(img_file, s[0]))
        scores_out.append(s[0])
    return scores_out
