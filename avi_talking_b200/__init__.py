"""avi_talking_b200 - B200-native (sm_100a) implementation of AVI-Talking's audio -> FLAME-vertex inference path.

Drop-in classes (same names / signatures / state_dict keys as the reference):
  avi_talking_b200.wav2vec.Wav2Vec2Model          <- models/lib/wav2vec.py
  avi_talking_b200.faceformer.Faceformer          <- models/faceformer_disentangle.py (FaceformerVert <- models/faceformer_vert.py)
  avi_talking_b200.flame.FLAME / FLAME_mediapipe  <- {gdl,inferno}/models/DecaFLAME.py ;  flame.lbs <- {gdl,inferno}/utils/lbs.py
All computation runs in libavi_b200.so (include/avi_b200.h); there is no CPU / eager-PyTorch fallback.
"""
__version__ = "0.1.0"
