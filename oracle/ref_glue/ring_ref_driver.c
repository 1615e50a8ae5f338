/* ORACLE / TEST INFRASTRUCTURE ONLY — links against the UNMODIFIED reference file
 * /root/reference/Core/Src/dsp_if.c (compiled in place by oracle/Makefile, never copied).
 *
 * Supplies the externals dsp_if.c expects from the firmware (HAL handle objects, the DMA
 * start call, the codec driver entry points — all no-ops on the host) and a few accessors so
 * a ctypes test can look at the ring state the reference keeps in file-scope globals
 * (dsp_if.c:32-35).
 */
#include <string.h>
#include "dsp_if.h"
#include "codec_if.h"

I2S_HandleTypeDef hi2s2;
I2C_HandleTypeDef hi2c3;

extern I2S_Buff_TypeDef i2s_buff;
extern DSP_Buff_TypeDef dsp_out_buff;
extern DSP_Buff_TypeDef dsp_in_buff;

HAL_StatusTypeDef HAL_I2SEx_TransmitReceive_DMA (I2S_HandleTypeDef *h, uint16_t *tx, uint16_t *rx, uint16_t n)
{ (void) h; (void) tx; (void) rx; (void) n; return HAL_OK; }
void Codec_Init (uint32_t f) { (void) f; }
void Codec_Set_RX (void) {}
void Codec_Set_TX (void) {}
uint8_t Codec_AF_Vol (uint8_t v) { return v; }
void Error_Handler (void) {}

/* the I2S callbacks live in dsp_if.c too */
void HAL_I2SEx_TxRxHalfCpltCallback (I2S_HandleTypeDef *hi2s);
void HAL_I2SEx_TxRxCpltCallback (I2S_HandleTypeDef *hi2s);

/* ---- accessors (ours) ---- */
uint32_t refring_fs (void)            { return USBD_AUDIO_FREQ; }
uint32_t refring_i2s_buff_size (void) { return I2S_BUFF_SIZE; }
uint32_t refring_i2s_half_size (void) { return I2S_BUFF_HALF_SIZE; }
uint32_t refring_dsp_buff_size (void) { return DSP_BUFF_SIZE; }
uint32_t refring_dsp_half_size (void) { return DSP_BUFF_HALF_SIZE; }

/* power-on state: the firmware's globals are zero-initialised .bss */
void refring_reset (void)
{
  memset (&i2s_buff, 0, sizeof i2s_buff);
  memset (&dsp_out_buff, 0, sizeof dsp_out_buff);
  memset (&dsp_in_buff, 0, sizeof dsp_in_buff);
}

/* which: 0 = dsp_in_buff (RX ring), 1 = dsp_out_buff (TX ring); out[3] = {enable, rd, wr} */
void refring_get_ptrs (int which, uint32_t *out)
{
  DSP_Buff_TypeDef *b = which ? &dsp_out_buff : &dsp_in_buff;
  out[0] = b->buff_enable; out[1] = b->rd_ptr; out[2] = b->wr_ptr;
}
void refring_get_iq (int which, int16_t *i, int16_t *q)
{
  DSP_Buff_TypeDef *b = which ? &dsp_out_buff : &dsp_in_buff;
  memcpy (i, b->i, sizeof b->i); memcpy (q, b->q, sizeof b->q);
}
uint16_t *refring_i2s_rx (void) { return i2s_buff.rx; }
uint16_t *refring_i2s_tx (void) { return i2s_buff.tx; }

/* one I2S DMA event: copy a codec block into the DMA half the ISR will read, fire the
 * reference callback for that half, hand back what it put in the TX half.
 * half: 0 = first half (HalfCplt), 1 = second half (Cplt). */
void refring_i2s_event (int half, const uint16_t *adc_block, uint16_t *dac_block)
{
  uint16_t *rx = i2s_buff.rx + (half ? I2S_BUFF_HALF_SIZE : 0);
  uint16_t *tx = i2s_buff.tx + (half ? I2S_BUFF_HALF_SIZE : 0);
  memcpy (rx, adc_block, I2S_BUFF_HALF_SIZE * sizeof (uint16_t));
  if (half) HAL_I2SEx_TxRxCpltCallback (&hi2s2); else HAL_I2SEx_TxRxHalfCpltCallback (&hi2s2);
  if (dac_block) memcpy (dac_block, tx, I2S_BUFF_HALF_SIZE * sizeof (uint16_t));
}
