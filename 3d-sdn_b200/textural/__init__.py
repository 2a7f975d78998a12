"""textural -- B200-native mirror of the hot operator of /root/reference/textural: the pix2pixHD GlobalGenerator forward
(BASELINE config 5).  Only `textural.models.networks.GlobalGenerator` lives here."""
