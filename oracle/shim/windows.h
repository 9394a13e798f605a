/* Win32 console stubs used by Scene::raytraceImage's Esc polling: never report an event. */
#ifndef MIRO_ORACLE_WINDOWS_STUB
#define MIRO_ORACLE_WINDOWS_STUB
typedef void* HANDLE; typedef unsigned long DWORD; typedef unsigned short WORD;
#define STD_INPUT_HANDLE 0
#define STD_OUTPUT_HANDLE 1
#define VK_ESCAPE 0x1B
struct KEY_EVENT_RECORD { WORD wVirtualKeyCode; };
struct INPUT_RECORD { union { KEY_EVENT_RECORD KeyEvent; } Event; };
static inline HANDLE GetStdHandle(int) { return 0; }
static inline int FlushConsoleInputBuffer(HANDLE) { return 1; }
static inline int PeekConsoleInput(HANDLE, INPUT_RECORD*, DWORD, DWORD* n) { *n = 0; return 1; }
static inline int ReadConsoleInput(HANDLE, INPUT_RECORD*, DWORD, DWORD* n) { *n = 0; return 1; }
#endif
