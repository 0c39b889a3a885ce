#include <stdlib.h>
static int imax(int a,int b){return a>b?a:b;}
/* last-row H and F of Gotoh DP of q (rows) vs t (cols); returns best */
int rows(const unsigned char* q,int lq,const unsigned char* t,int lt,int ma,int mi,int gi,int ge,int* Hout,int* Fout){
  int* H=calloc(lt+1,sizeof(int)); int* F=calloc(lt+1,sizeof(int)); int best=0;
  for(int i=1;i<=lq;i++){ int e=0,hl=0,hd=0; for(int j=1;j<=lt;j++){ e=imax(e-ge,hl-gi); int f=imax(F[j]-ge,H[j]-gi); int h=hd+(q[i-1]==t[j-1]?ma:mi); if(e>h)h=e; if(f>h)h=f; if(h<0)h=0; hd=H[j]; H[j]=h; F[j]=f; hl=h; if(h>best)best=h; } }
  for(int j=0;j<=lt;j++){Hout[j]=H[j];Fout[j]=F[j];} free(H);free(F); return best; }
