"""Constants of Tracks T/V (reference: constants.py:4-17, config.py:5-32).  Hyper-parameters and the RAF-DB class
list are facts of the workload and are kept identical; the prompt texts below are this repo's own wording (any
`{emotion: [prompts]}` dict can be passed to CLIPAdapter(emotion_descriptions=...), including the reference's)."""
import torch

DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")

# Track T (constants.py)
BATCH_SIZE = 32
MODEL_NAME = "openai/clip-vit-large-patch14"
BOTTLENECK_DIM = 64
LEARNING_RATE = 3e-4
NUM_EPOCHS = 5
ALPHA = 0.2
BETA = 0.2

# Track V (config.py)
CLIP_MODEL_NAME = "openai/clip-vit-large-patch14"
V_BATCH_SIZE = 4
V_BOTTLENECK_DIM = 192
GAMMA = 0.3
SEED = 42

EMOTIONS = ["angry", "disgust", "fear", "happy", "neutral", "sad", "surprise"]

_TEMPLATES = [
    "a photo of a face that looks {}",
    "a close-up portrait of a {} person",
    "a facial expression showing that the person is {}",
    "someone whose face is clearly {}",
    "a cropped face image, the emotion is {}",
]
_WORDS = {"angry": "angry", "disgust": "disgusted", "fear": "afraid", "happy": "happy", "neutral": "neutral",
          "sad": "sad", "surprise": "surprised"}


def get_emotion_descriptions():
    """Five prompts per RAF-DB class (same structure as reference constants.py:20-75: 7 classes x 5 prompts)."""
    return {e: [t.format(_WORDS[e]) for t in _TEMPLATES] for e in EMOTIONS}

# pixel normalisation constants (RGB): process_video.py:24 uses ImageNet's; CLIPImageProcessor uses CLIP's
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
CLIP_MEAN, CLIP_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)
