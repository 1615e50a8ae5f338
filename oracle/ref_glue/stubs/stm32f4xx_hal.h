/* ORACLE / TEST INFRASTRUCTURE ONLY — never shipped, never on the product path.
 *
 * Minimal stand-in for the ST HAL umbrella header so that the UNMODIFIED
 * /root/reference/Core/Src/dsp_if.c compiles on the x86 host.
 * Only what dsp_if.c / dsp_if.h / main.h / codec_if.h reference is declared:
 *   I2S_HandleTypeDef, I2C_HandleTypeDef, HAL_StatusTypeDef and the prototype of
 *   HAL_I2SEx_TransmitReceive_DMA (reference: stm32f4xx_hal_i2s_ex.h:145).
 */
#ifndef SLO_STUB_STM32F4XX_HAL_H
#define SLO_STUB_STM32F4XX_HAL_H
#include <stdint.h>
typedef enum { HAL_OK = 0, HAL_ERROR = 1, HAL_BUSY = 2, HAL_TIMEOUT = 3 } HAL_StatusTypeDef;
typedef struct { int unused; } I2S_HandleTypeDef;
typedef struct { int unused; } I2C_HandleTypeDef;
HAL_StatusTypeDef HAL_I2SEx_TransmitReceive_DMA (I2S_HandleTypeDef *hi2s, uint16_t *pTxData,
                                                 uint16_t *pRxData, uint16_t Size);
#endif
