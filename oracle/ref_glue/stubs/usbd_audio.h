/* ORACLE / TEST INFRASTRUCTURE ONLY.
 * Stand-in for USB_DEVICE/Class/usbd_audio.h: dsp_if.h only needs the two geometry
 * macros (reference: usbd_audio.h:46 and :53). USBD_AUDIO_FREQ is overridable with -D
 * so the same unmodified dsp_if.c can be built at 48/96/192 kHz. */
#ifndef SLO_STUB_USBD_AUDIO_H
#define SLO_STUB_USBD_AUDIO_H
#ifndef USBD_AUDIO_FREQ
#define USBD_AUDIO_FREQ 96000U
#endif
#define AUDIO_OUT_PACKET_NUM 2U
#endif
